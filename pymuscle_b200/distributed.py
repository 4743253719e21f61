"""
Muscle batches sharded over the GPUs of one box (one process per GPU).

Muscles never read each other's state (the reference keeps one pool and one
fibers object per Muscle, pymuscle/muscle.py:50-51, and callers loop over
separate instances, examples/envs/muscled_hopper.py:33-34), so the batch axis
is cut into contiguous blocks of rows, one per rank, and every rank advances
its own rows with the same kernels as a single-GPU MuscleBatch.  There is NO
collective on the data path.  The only exchange is the optional gather of
per-muscle totals (``gather=True`` / ``gather_totals``): one all-gather over
``torch.distributed`` (NCCL over NVLink on the GPUs, gloo in the CPU tests),
issued per step or per fused window, optionally asynchronously so that it
overlaps the next window -- or, for fused windows, ``gather="peer"``: the kernel
itself stores every total into all GPUs' gather buffers (symmetric memory, plain
stores over NVLink), so the gather overlaps the arithmetic step by step and no
collective kernel runs at all.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple, Union

import torch
import torch.distributed as dist

from .tables import MuscleSpec


def shard_bounds(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Rows [lo, hi) owned by ``rank``: contiguous blocks, the first ``total % world``
    ranks hold one row more."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} of world {world}")
    if total < 0:
        raise ValueError("total must be non-negative")
    base, extra = divmod(int(total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(total: int, world: int):
    return [shard_bounds(total, r, world)[1] - shard_bounds(total, r, world)[0] for r in range(world)]


class _Handle:
    """An all-gather in flight: ``wait()`` returns the gathered tensor."""

    def __init__(self, work, pieces, widths, tail):
        self._work, self._pieces, self._widths, self._tail = work, pieces, widths, tail

    def wait(self) -> torch.Tensor:
        if self._work is not None:
            self._work.wait()
            self._work = None
        t = torch.cat([p[..., :w] for p, w in zip(self._pieces, self._widths)], dim=-1) \
            if len(self._pieces) > 1 else self._pieces[0]
        return t.reshape(*t.shape[:-1], *self._tail) if self._tail else t

    def result(self, async_op: bool):
        return self if async_op else self.wait()


class ShardedMuscleBatch:
    """``total_rows`` rows of ``spec`` spread over the ranks of ``group``.

    Every rank constructs this with the same arguments.  ``step`` / ``step_many`` accept
    either this rank's slice of the excitations or the global tensor (which is then
    sliced); they return this rank's totals, or the totals of all rows when
    ``gather=True``.

    engine_cls   the per-rank engine (default: MuscleBatch, CUDA only).  The CPU tests
                 inject an oracle-backed engine to exercise the sharding logic over gloo.
    """

    def __init__(self, spec: Union[MuscleSpec, Sequence[MuscleSpec]], total_rows: int,
                 device=None, precision="f32", group: Optional[dist.ProcessGroup] = None,
                 engine_cls=None, **engine_kwargs):
        self.group = group
        if dist.is_available() and dist.is_initialized():
            self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        else:
            self.rank, self.world = 0, 1
        self.total_rows = int(total_rows)
        if self.total_rows < self.world:
            raise ValueError(f"{self.total_rows} rows cannot be spread over {self.world} ranks")
        self.row_lo, self.row_hi = shard_bounds(self.total_rows, self.rank, self.world)
        self.sizes = shard_sizes(self.total_rows, self.world)
        if engine_cls is None:
            from .engine import MuscleBatch
            engine_cls = MuscleBatch
        self.local = engine_cls(spec, self.row_hi - self.row_lo, device, precision, **engine_kwargs)
        self.S = self.local.S

    # ---- shapes -----------------------------------------------------------------------
    @property
    def local_rows(self) -> int:
        return self.row_hi - self.row_lo

    @property
    def n_muscles_local(self) -> int:
        return self.local_rows * self.S

    @property
    def n_muscles_total(self) -> int:
        return self.total_rows * self.S

    def _local_slice(self, x, per_step: bool, global_input: Optional[bool]):
        """This rank's part of an excitation argument.  The muscle axis (axis 1 of a
        [K, ...] window, else axis 0) may hold this rank's rows or all rows, as [rows],
        [rows * S] or [rows * L]; ``global_input`` forces the interpretation, None guesses
        (a size that fits the local shard is taken as local)."""
        ax = 1 if per_step else 0
        if self.world == 1 or not isinstance(x, torch.Tensor) or x.dim() <= ax or global_input is False:
            return x
        size, S, L = int(x.shape[ax]), self.S, self.local.L
        flat = x.dim() == ax + 1
        if global_input is None and (size == self.local_rows or (flat and size in (self.local_rows * S, self.local_rows * L))):
            return x
        for per_row in ((1, S, L) if flat else (1,)):
            if size == self.total_rows * per_row:
                return x.narrow(ax, self.row_lo * per_row, self.local_rows * per_row)
        if global_input:
            raise AssertionError(f"excitation of shape {tuple(x.shape)} does not cover {self.total_rows} rows")
        return x

    # ---- stepping ---------------------------------------------------------------------
    def step(self, excitation, step_size: float, gather: bool = False,
             global_input: Optional[bool] = None, **kw):
        out = self.local.step(self._local_slice(excitation, False, global_input), step_size, **kw)
        return self.gather_totals(out) if gather else out

    def step_many(self, excitations, step_size: float, gather=False,
                  global_input: Optional[bool] = None, **kw):
        """``gather``: False (local totals), True (all-gather over torch.distributed) or "peer"
        (the kernel stores into all GPUs' buffers; needs ``enable_peer_gather``)."""
        if gather == "peer":
            return self.step_many_peer_gather(excitations, step_size, global_input)
        per_step = kw.get("steps") is None                   # [M] + steps=K is the constant form
        out = self.local.step_many(self._local_slice(excitations, per_step, global_input), step_size, **kw)
        if gather:
            tot = out[0] if isinstance(out, tuple) else out
            g = self.gather_totals(tot)
            return (g,) + tuple(out[1:]) if isinstance(out, tuple) else g
        return out

    # ---- gather fused into the kernel: peer stores over NVLink -------------------------------
    # Gather buffer (identical layout on every rank): one region per SOURCE rank r, laid out so that what one warp
    # stores per 32-step tile is 256 contiguous bytes of NVLink traffic:
    #   FP32 mode (one warp per muscle PAIR): pair-major    region_r[pair][k][2],  pair = local muscle >> 1, k < max_steps
    #   FP64 mode (one warp per muscle):      muscle-major  region_r[muscle][k]    (2 * pairs muscle slots)
    def enable_peer_gather(self, max_steps: int, multicast: Optional[bool] = None, layout: str = "auto"):
        """Collective.  Allocates two symmetric gather buffers per rank (double-buffered between
        consecutive windows) and exchanges their mappings.  ``multicast``: store through the NVLink multicast
        mapping of the buffers (NVLS: one store per total leaves the GPU, the switch replicates it) instead of
        one store per peer; None = use it when the platform offers it."""
        import torch.distributed._symmetric_memory as symm_mem
        if self.S != 1:
            raise ValueError("the fused gather covers uniform batches")
        group = self.group if self.group is not None else dist.group.WORLD
        ks = int(max_steps)
        pairs = [(m + 1) // 2 for m in self.sizes]
        offs = [0]
        for p in pairs:
            offs.append(offs[-1] + p * ks * 2)
        self._pg = {"steps": ks, "i": 0, "buf": [], "hdl": [], "pairs": pairs, "offs": offs}
        for _ in range(2):
            t = symm_mem.empty(offs[-1], dtype=self.local.dtype, device=self.local.device)
            self._pg["buf"].append(t)
            self._pg["hdl"].append(symm_mem.rendezvous(t, group))
        have_mc = all(int(getattr(h, "multicast_ptr", 0) or 0) != 0 for h in self._pg["hdl"])
        if multicast and not have_mc:
            raise RuntimeError("NVLink multicast is not available for this group")
        self._pg["multicast"] = have_mc if multicast is None else bool(multicast)
        if layout not in ("auto", "pair", "muscle"):
            raise ValueError("layout must be 'auto', 'pair' or 'muscle'")
        self._pg["layout"] = layout                             # "auto": muscle-major in FP64 mode, pair-major in FP32 mode

    def step_many_peer_gather(self, excitations: torch.Tensor, step_size: float,
                              global_input: Optional[bool] = None, dense: bool = True):
        """A fused window whose totals land in EVERY rank's gather buffer through the kernel's own
        stores.  ``dense=True`` returns a contiguous [K, total muscles] tensor (one local re-layout
        pass); ``dense=False`` returns the per-source-rank views [K, muscles_r] (strided, no copy),
        valid until the window after next.  Consume the result on the current stream before
        launching further windows."""
        pg = getattr(self, "_pg", None)
        if pg is None:
            raise RuntimeError("call enable_peer_gather(max_steps) on every rank first")
        e = self._local_slice(excitations, True, global_input)
        K, ks = int(e.shape[0]), pg["steps"]
        if K > ks:
            raise ValueError(f"window of {K} steps exceeds enable_peer_gather({ks})")
        b = pg["i"] & 1
        pg["i"] += 1
        buf, hdl = pg["buf"][b], pg["hdl"][b]
        esz = buf.element_size()
        off = pg["offs"][self.rank] * esz
        mc = pg.get("multicast", False)
        if mc:                                                  # one multicast address: the switch fans the stores out
            peers = [int(hdl.multicast_ptr) + off]
        else:
            peers = [hdl.get_buffer(r, (buf.numel(),), buf.dtype).data_ptr() + off
                     for r in range(self.world) if r != self.rank]
        local = torch.empty(K, self.n_muscles_local, dtype=buf.dtype, device=buf.device)
        muscle_major = pg.get("layout", "auto") == "muscle" or (pg.get("layout", "auto") == "auto" and buf.dtype == torch.float64)
        if muscle_major:
            self.local.step_many_gather(e, step_size, local, peers, 1, 2 * ks, ks, multicast=mc)
        else:
            self.local.step_many_gather(e, step_size, local, peers, 2, 2 * ks, 1, multicast=mc)
        hdl.barrier()                                           # every rank's stores have landed everywhere
        views = []
        for r in range(self.world):
            if r == self.rank:
                views.append(local)
                continue
            if muscle_major:
                reg = buf[pg["offs"][r]: pg["offs"][r + 1]].view(2 * pg["pairs"][r], ks)
                v = reg[: self.sizes[r], :K].permute(1, 0)                        # [K, muscles_r], strided view
                views.append(v.contiguous() if dense else v)
                continue
            reg = buf[pg["offs"][r]: pg["offs"][r + 1]].view(pg["pairs"][r], ks, 2)
            views.append(reg[:, :K].permute(1, 0, 2).reshape(K, -1)[:, : self.sizes[r]] if dense
                         else reg[:, :K].permute(1, 0, 2))
        if dense:
            return torch.cat(views, dim=1)
        return views

    # ---- the one exchange: all ranks' per-muscle totals ---------------------------------
    def gather_totals(self, local: torch.Tensor, async_op: bool = False):
        """All-gather along the muscle axis.  ``local`` is [..., local_rows * S] (what ``step`` of
        a uniform batch and ``step_many`` return) or [..., local_rows, S] (``step`` of a ragged
        batch); the result has the same form with ``total_rows`` in place of ``local_rows``, on
        every rank (a handle with ``wait()`` when ``async_op``).  Shards of unequal size are
        padded to the largest one for the wire."""
        S = self.S
        two_d = S > 1 and local.dim() >= 2 and tuple(local.shape[-2:]) == (self.local_rows, S)
        flat = local.reshape(*local.shape[:-2], -1) if two_d else local
        if flat.shape[-1] != self.local_rows * S:
            raise AssertionError(
                f"expected {self.local_rows * S} local totals on the last axis, got {tuple(local.shape)}")
        tail = (self.total_rows, S) if two_d else None
        if self.world == 1:
            return _Handle(None, [flat], [flat.shape[-1]], tail).result(async_op)
        widths = [s * S for s in self.sizes]
        send = flat.contiguous()
        if send.shape[-1] != max(widths):
            pad = send.new_zeros(*send.shape[:-1], max(widths))
            pad[..., : send.shape[-1]] = send
            send = pad
        pieces = [torch.empty_like(send) for _ in range(self.world)]
        work = dist.all_gather(pieces, send, group=self.group, async_op=True)
        return _Handle(work, pieces, widths, tail).result(async_op)

    # ---- observers / state ----------------------------------------------------------------
    def get_peripheral_fatigue(self, gather: bool = False):
        out = self.local.get_peripheral_fatigue()
        return self.gather_totals(out) if gather else out

    def state_dict(self) -> dict:
        """This rank's shard of the state plus where it sits in the global batch."""
        sd = dict(self.local.state_dict())
        sd["row_range"] = (self.row_lo, self.row_hi)
        sd["total_rows"] = self.total_rows
        return sd

    def load_state_dict(self, state: dict):
        if tuple(state.get("row_range", (self.row_lo, self.row_hi))) != (self.row_lo, self.row_hi):
            raise ValueError("state_dict belongs to a different shard")
        self.local.load_state_dict({k: v for k, v in state.items() if k not in ("row_range", "total_rows")})
