"""
The N>1 path on CPU: world_size-2 (and 3) process groups over gloo.

ShardedMuscleBatch only partitions rows, slices excitations and gathers totals;
the per-rank arithmetic is whatever engine it is given.  Here the engine is a
small torch-CPU wrapper around the NumPy oracle (tests may use the oracle), so
the sharded run can be compared with an unsharded oracle run of the same batch.
On the GPU box the engine is MuscleBatch and the backend NCCL; the logic under
test is identical.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.pymuscle_oracle import OracleMuscles
from pymuscle_b200.distributed import ShardedMuscleBatch, shard_bounds, shard_sizes
from pymuscle_b200.tables import MuscleSpec


class OracleEngine:
    """The slice of MuscleBatch's interface that ShardedMuscleBatch uses, on the CPU."""

    def __init__(self, spec, batch, device=None, precision="f64"):
        assert isinstance(spec, MuscleSpec)
        kind = "standard" if spec.fibers.recovery else "potvin"
        self.o = OracleMuscles(spec.n, batch, kind)
        self.rows, self.S, self.L = batch, 1, spec.n

    def step(self, exc, dt):
        return torch.from_numpy(self.o.step(exc.numpy(), dt).total.copy())

    def step_many(self, exc, dt, steps=None):
        e = exc.numpy()
        if steps is not None:
            e = np.broadcast_to(e, (steps, self.rows))
        return torch.from_numpy(np.stack([self.o.step(e[k], dt).total for k in range(e.shape[0])]))

    def get_peripheral_fatigue(self):
        return torch.from_numpy(self.o.peripheral_fatigue())

    def state_dict(self):
        return {"current_peak_forces": torch.from_numpy(self.o.cpf.copy()),
                "recruitment_durations": torch.from_numpy(self.o.dur.copy())}

    def load_state_dict(self, sd):
        self.o.cpf = sd["current_peak_forces"].numpy().copy()
        self.o.dur = sd["recruitment_durations"].numpy().copy()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total_rows, kind, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        spec = MuscleSpec.potvin(24) if kind == "potvin" else MuscleSpec.standard()
        scale = 67.0 if kind == "potvin" else 1.0
        rng = np.random.default_rng(5)
        K, dt = 6, 0.02
        exc = torch.from_numpy(scale * rng.random((K + 2, total_rows)))          # the same on every rank
        sb = ShardedMuscleBatch(spec, total_rows, precision="f64", engine_cls=OracleEngine)
        lo, hi = shard_bounds(total_rows, rank, world)
        assert (sb.row_lo, sb.row_hi) == (lo, hi) and sb.local.rows == hi - lo

        # single steps: global input sliced by the wrapper, totals gathered on every rank
        got0 = sb.step(exc[0], dt, gather=True, global_input=True)
        # local input, local output, then an explicit asynchronous gather
        loc1 = sb.step(exc[1, lo:hi], dt)
        assert tuple(loc1.shape) == (hi - lo,)
        got1 = sb.gather_totals(loc1, async_op=True).wait()
        # a fused window [K, rows], gathered along the muscle axis
        gotw = sb.step_many(exc[2:], dt, gather=True, global_input=True)
        fat = sb.get_peripheral_fatigue(gather=True) if kind == "standard" else None

        # unsharded run of the same batch
        ref = OracleEngine(spec, total_rows)
        want0, want1 = ref.step(exc[0], dt), ref.step(exc[1], dt)
        wantw = ref.step_many(exc[2:], dt)
        assert torch.equal(got0, want0) and torch.equal(got1, want1)
        assert tuple(gotw.shape) == (K, total_rows) and torch.equal(gotw, wantw)
        if fat is not None:
            assert torch.equal(fat, ref.get_peripheral_fatigue())

        # the shard's state is the matching block of the unsharded state, and reloads
        sd = sb.state_dict()
        assert sd["row_range"] == (lo, hi) and sd["total_rows"] == total_rows
        assert torch.equal(sd["current_peak_forces"], ref.state_dict()["current_peak_forces"][lo:hi])
        sb.load_state_dict(sd)
        with pytest.raises(ValueError):
            sb.load_state_dict(dict(sd, row_range=(lo + 1, hi + 1)))
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,total_rows,kind", [(2, 8, "potvin"), (2, 7, "standard"), (3, 8, "potvin")])
def test_sharded_batch_over_gloo(world, total_rows, kind):
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total_rows, kind, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert dict(ret) == {r: "ok" for r in range(world)}


def test_shard_bounds_cover_the_batch_exactly():
    for total in (1, 7, 8, 65536, 262144, 4096):
        for world in (1, 2, 3, 4, 8):
            if total < world:
                continue
            spans = [shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = shard_sizes(total, world)
            assert max(sizes) - min(sizes) <= 1 and sum(sizes) == total
    with pytest.raises(ValueError):
        shard_bounds(8, 2, 2)


def test_single_process_is_a_plain_batch():
    sb = ShardedMuscleBatch(MuscleSpec.potvin(12), 5, precision="f64", engine_cls=OracleEngine)
    assert (sb.world, sb.rank, sb.row_lo, sb.row_hi) == (1, 0, 0, 5)
    e = torch.full((5,), 40.0, dtype=torch.float64)
    t = sb.step(e, 0.02, gather=True)
    assert tuple(t.shape) == (5,)
