"""
The vector-environment step (mb_step_env through MuscleBatch.step_env / MuscleGroupEnv):
outputs, Hill-type scaling and the fused peripheral-fatigue readout against the oracle's
per-muscle loop -- the reference's `[muscle.step(a[i], dt) for i, muscle in enumerate(muscles)]`
(examples/envs/muscled_hopper.py:33-34) followed by get_peripheral_fatigue (muscle.py:248-256).
"""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle.pymuscle_oracle import OracleMuscles
from pymuscle_b200 import MuscleBatch, MuscleGroupEnv, MuscleSpec
from test_gpu_parity import close

pytestmark = pytest.mark.gpu

FORCES = [48.0, 48.5, 35.0, 21.0]          # a hopper-like leg: four muscles of different size


@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_env_step_matches_the_per_muscle_loop(cuda_device, precision):
    specs = [MuscleSpec.standard(f) for f in FORCES]
    agents, dt, frames = 37, 0.02, 25
    env = MuscleGroupEnv(specs, agents, cuda_device, precision)
    oras = [OracleMuscles.standard(f, m=agents) for f in FORCES]
    rng = np.random.default_rng(11)
    for k in range(frames):
        a = rng.random((agents, len(specs))).astype(np.float32)
        out, fat = env.step(torch.from_numpy(a), dt)
        for s, o in enumerate(oras):
            want = o.step(a[:, s].astype(np.float64), dt)
            close(out[:, s].cpu().numpy(), want.total, precision, 1.0, f"outputs, muscle {s}, frame {k}")
            close(fat[:, s].cpu().numpy(), o.peripheral_fatigue(), precision, 1.0, f"fatigue, muscle {s}, frame {k}")
    assert fat.dtype == torch.float64 and bool((fat >= 0).all()) and bool((fat < 1).all())


def test_env_step_is_one_launch_and_equals_step_plus_fatigue(cuda_device):
    """The fused epilogue gives the same numbers as the two-kernel path, in one launch."""
    g = golden("arm")
    specs = [MuscleSpec.standard(max_force=float(f)) for f in g["max_forces"]]
    a = MuscleBatch(specs, 16, cuda_device, "f32", fresh_outputs=False)
    b = MuscleBatch(specs, 16, cuda_device, "f32", fresh_outputs=False)
    x = torch.rand(16, 30, device=cuda_device)
    for _ in range(5):
        n0 = a.launches
        out, fat = a.step_env(x, 0.02)
        assert a.launches == n0 + 1
        want = b.step(x, 0.02)
        assert torch.equal(out, want)
        np.testing.assert_allclose(fat.cpu().numpy(), b.get_peripheral_fatigue().cpu().numpy(), rtol=0, atol=2e-7)
    assert torch.equal(a.cpf, b.cpf)


def test_hill_type_gain_scales_the_outputs(cuda_device):
    spec = MuscleSpec.standard()
    agents = 64
    rest = torch.full((1,), 0.3)
    plain = MuscleGroupEnv(spec, agents, cuda_device, "f32")
    hill = MuscleGroupEnv(spec, agents, cuda_device, "f32", rest_lengths=rest)
    from pymuscle_b200 import (contractile_element_force_length_curve as fl_curve,
                               contractile_element_force_velocity_curve as fv_curve)
    rng = np.random.default_rng(2)
    prev = None
    for k in range(6):
        act = torch.from_numpy(rng.random((agents, 1)).astype(np.float32))
        cur = torch.from_numpy((0.3 * (0.8 + 0.5 * rng.random((agents, 1)))).astype(np.float32))
        o0, f0 = plain.step(act, 0.02)
        o1, f1 = hill.step(act, 0.02, lengths=cur)
        p = cur if prev is None else prev
        gain = fl_curve(0.3, cur.double().numpy()) * fv_curve(0.3, cur.double().numpy(), p.double().numpy(), 0.02)
        # float32 evaluation of 1.8 - 1.8/(1 + e) cancels to 0 for fast lengthening: absolute term
        np.testing.assert_allclose(o1.cpu().numpy(), o0.cpu().numpy() * gain, rtol=2e-6, atol=1e-6)
        assert torch.equal(f0, f1)                             # the gain does not touch the fatigue state
        prev = cur
    with pytest.raises(ValueError):
        plain.step(act, 0.02, lengths=cur)


def test_env_step_falls_back_on_an_odd_row_length(cuda_device):
    """L = 33 units: the streaming kernel declines, the generic kernels serve the same call."""
    spec = MuscleSpec.potvin(33)
    b = MuscleBatch(spec, 9, cuda_device, "f64")
    o = OracleMuscles(33, 9, "potvin")
    x = np.linspace(5.0, 60.0, 9)
    out, fat = b.step_env(torch.from_numpy(x), 0.05)
    close(out.cpu().numpy(), o.step(x, 0.05).total, "f64", 500.0, "totals")
    # the readout is sum(peak - capacity) / divisor (include/muscle_b200.h): 1 - sum(capacity)/max_arb_output of a
    # StandardMuscle (muscle.py:248-256, max_arb_output = sum(peak)); a Potvin muscle has divisor 1 and no reference method
    want_f = np.array([(o.t.ptf - o.cpf[i]).sum() / 1.0 for i in range(9)])
    np.testing.assert_allclose(fat.cpu().numpy(), want_f, rtol=1e-9)


@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_graphed_env_step_replays_the_same_step(cuda_device, precision):
    """CUDA-graph replay of the env step == the eager call, bit for bit, state included."""
    specs = [MuscleSpec.standard(f) for f in (20.0, 32.0, 45.0, 32.0)]
    agents = 37
    eager, graphed = MuscleGroupEnv(specs, agents, cuda_device, precision), MuscleGroupEnv(specs, agents, cuda_device, precision)
    f = graphed.graphed_step(0.02, with_gain=True)
    g = torch.Generator(device=cuda_device).manual_seed(5)
    for k in range(25):
        a = torch.rand(agents, 4, device=cuda_device, generator=g)
        gain = 0.5 + torch.rand(agents, 4, device=cuda_device, generator=g)
        out_e, fat_e = eager.batch.step_env(a.to(eager.dtype), 0.02, gain=gain.to(eager.dtype))
        out_g, fat_g = f(a, gain)
        assert torch.equal(out_e.view(agents, 4), out_g) and torch.equal(fat_e.view(agents, 4), fat_g), f"step {k}"
    assert torch.equal(eager.batch.cpf, graphed.batch.cpf)
    assert torch.equal(eager.batch.current_forces, graphed.batch.current_forces)


@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_rows_of_small_muscles_share_column_tiles(cuda_device, precision):
    """Rows made of many SMALL muscles: the streaming kernel packs consecutive muscles into one column
    tile (every consumer warp inside one muscle).  Twelve muscles of 9..264 units (957 per row, odd):
    packed groups with muscles narrower than a warp, single-muscle groups, and muscles wider than a tile;
    per-muscle and per-unit inputs, the fatigue epilogue, 21 rows (partial chunk)."""
    forces = [3.0, 5.0, 8.0, 10.0, 12.0, 16.0, 20.0, 32.0, 45.0, 32.0, 70.0, 5.0]
    specs = [MuscleSpec.standard(f) for f in forces]
    assert [s.n for s in specs] == [9, 17, 28, 36, 44, 59, 74, 120, 169, 120, 264, 17]
    rows, S, rng = 21, len(specs), np.random.default_rng(12)
    offs = np.cumsum([0] + [s.n for s in specs])
    oras, oras_u = [OracleMuscles.standard(f, m=rows) for f in forces], [OracleMuscles.standard(f, m=rows) for f in forces]
    b, bu = MuscleBatch(specs, rows, cuda_device, precision), MuscleBatch(specs, rows, cuda_device, precision)
    for k in range(15):
        e = rng.random((rows, S)).astype(np.float32)
        eu = rng.random((rows, int(offs[-1]))).astype(np.float32)
        if k % 2:
            got, fat = b.step_env(torch.from_numpy(e).to(cuda_device, b.dtype), 0.1)
        else:
            got, fat = b.step(torch.from_numpy(e).to(cuda_device, b.dtype), 0.1), None
        got_u = bu.step(torch.from_numpy(eu).to(cuda_device, b.dtype), 0.1)
        for j, s in enumerate(specs):
            ptf = s.tables()["ptf"]
            want = oras[j].step(e[:, j].astype(np.float64), 0.1)
            want_u = oras_u[j].step(eu[:, offs[j]:offs[j + 1]].astype(np.float64), 0.1)
            close(got[:, j].cpu().numpy(), want.total, precision, 1.0, f"totals muscle {j} step {k}")
            close(got_u[:, j].cpu().numpy(), want_u.total, precision, 1.0, f"per-unit-input totals muscle {j} step {k}")
            close(b.current_forces[:, offs[j]:offs[j + 1]].cpu().numpy(), want.forces, precision, ptf, f"forces {j}/{k}")
            close(b.cpf[:, offs[j]:offs[j + 1]].cpu().numpy(), oras[j].cpf, precision, ptf, f"capacities {j}/{k}")
            close(bu.cpf[:, offs[j]:offs[j + 1]].cpu().numpy(), oras_u[j].cpf, precision, ptf, f"per-unit-input capacities {j}/{k}")
            if fat is not None:
                close(fat[:, j].cpu().numpy(), oras[j].peripheral_fatigue(), precision, 1.0, f"fatigue {j}/{k}")
