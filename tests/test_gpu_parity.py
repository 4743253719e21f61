"""
GPU parity: the CUDA path (through the C ABI, driven by MuscleBatch) against
golden trajectories of the unmodified reference and against the oracle.

Bars (BASELINE.json north_star): recruitment masks bit-exact; per-unit forces,
capacities and totals within rtol 1e-9 (FP64 mode) / 1e-4 (FP32 mode) over
10,000 steps.  Units do reach capacity exactly 0 and the crossing step can move
by one between implementations, so every comparison carries a
small absolute term next to rtol: 1e-12 x scale (FP64) / 1e-6 x scale (FP32),
scale = the unit's peak twitch force (per-unit values) or the muscle's maximal
output (totals).
"""
import numpy as np
import pytest
import torch

from conftest import golden, trajectory_names, traj_kwargs
from oracle.pymuscle_oracle import OracleMuscles, PoolHyper, UnitTables, recruitment_mask
from pymuscle_b200 import MuscleBatch, MuscleSpec
from pymuscle_b200 import _native as nat

pytestmark = pytest.mark.gpu

RTOL = {"f32": 1e-4, "f64": 1e-9}
ATOL_REL = {"f32": 1e-6, "f64": 1e-12}
STD_FORCE_FOR_N = {120: 32.0, 340: 90.0, 2000: 527.1}


def make_spec(kind, n, kw):
    if kind == "potvin":
        return MuscleSpec.potvin(n, **kw)
    spec = MuscleSpec.standard(STD_FORCE_FOR_N[n], **kw)
    assert spec.n == n
    return spec


def close(got, want, precision, scale, what):
    got = np.asarray(got, dtype=np.float64)
    err = np.abs(got - want)
    bound = RTOL[precision] * np.abs(want) + ATOL_REL[precision] * scale
    bad = err > bound
    assert not bad.any(), (f"{what}: {bad.sum()} of {bad.size} outside tolerance; worst excess "
                           f"{(err - bound).max():.3e}, worst rel {np.max(err / np.maximum(np.abs(want), 1e-300)):.3e}")


def total_scale(spec):
    return float(np.sum(spec.tables()["ptf"])) / spec.output_divisor


@pytest.mark.parametrize("precision", ["f32", "f64"])
@pytest.mark.parametrize("mode", ["single_step", "fused"])
@pytest.mark.parametrize("name", trajectory_names())
def test_trajectory(cuda_device, name, mode, precision):
    g = golden("traj_" + name)
    exc, dt, n, kind = g["exc"], float(g["dt"]), int(g["n"]), str(g["kind"])
    per_unit = bool(g["per_unit"])
    if per_unit and mode == "fused":
        pytest.skip("per-unit excitation vectors are a single-step feature")
    K, m = exc.shape[:2]
    spec = make_spec(kind, n, traj_kwargs(g))
    ptf = spec.tables()["ptf"]
    b = MuscleBatch(spec, m, cuda_device, precision)
    x = torch.from_numpy(exc).to(cuda_device, b.dtype)
    totals = torch.empty(K, m, dtype=b.dtype, device=cuda_device)
    prev = 0
    cps = [int(c) for c in g["checkpoints"]]
    for i, cp in enumerate(cps + ([K - 1] if cps[-1] != K - 1 else [])):
        if mode == "single_step":
            for k in range(prev, cp + 1):
                totals[k] = b.step(x[k], dt)
        else:
            totals[prev:cp + 1] = b.step_many(x[prev:cp + 1], dt)
        prev = cp + 1
        if i < len(cps):
            close(b.current_forces.cpu().numpy(), g["forces"][i], precision, ptf, f"forces@{cp}")
            close(b.cpf.cpu().numpy(), g["cpf"][i], precision, ptf, f"capacities@{cp}")
            np.testing.assert_allclose(b.dur.cpu().numpy(), g["dur"][i], rtol=1e-9, atol=1e-12)
    close(totals.cpu().numpy(), g["totals"], precision, total_scale(spec), "totals")


@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_recruitment_mask_probes(cuda_device, precision):
    """Inputs within a few ulps of every threshold: the mask must be the reference's, bit for bit."""
    g = golden("mask_probes")
    cases = [("probes32", "mask32", MuscleSpec.potvin(120)), ("xprobes32", "xmask32", MuscleSpec.standard())]
    if precision == "f64":
        cases.append(("probes64", "mask64", MuscleSpec.potvin(120)))
    for pk, mk, spec in cases:
        probes = g[pk]
        want = np.unpackbits(g[mk], axis=1)[:, :120]
        b = MuscleBatch(spec, len(probes), cuda_device, precision)
        mask = torch.empty(len(probes), 120, dtype=torch.uint8, device=cuda_device)
        b.step(torch.from_numpy(probes.astype(np.float64)).to(cuda_device, b.dtype), 0.02, mask_out=mask)
        assert np.array_equal(mask.cpu().numpy(), want), pk
        # per-unit form of the same inputs
        b = MuscleBatch(spec, len(probes), cuda_device, precision)
        xu = torch.from_numpy(np.repeat(probes.astype(np.float64)[:, None], 120, axis=1)).to(cuda_device, b.dtype)
        b.step(xu, 0.02, mask_out=mask)
        assert np.array_equal(mask.cpu().numpy(), want), pk + " per-unit"


@pytest.mark.parametrize("precision", ["f32", "f64"])
@pytest.mark.parametrize("kind", ["potvin", "standard"])
def test_masks_random_excitation_fused_and_single(cuda_device, kind, precision):
    rng = np.random.default_rng(11)
    m, K, n = 37, 64, 120
    spec = MuscleSpec.potvin(n) if kind == "potvin" else MuscleSpec.standard()
    exc = (rng.random((K, m)) * (67.0 if kind == "potvin" else 1.0)).astype(np.float32)
    t, h = UnitTables.build(n), PoolHyper()
    e64 = exc.astype(np.float64)[:, :, None] * spec.input_scale
    want = recruitment_mask(np.broadcast_to(e64, (K, m, n)), t, h).astype(np.uint8)
    b = MuscleBatch(spec, m, cuda_device, precision)
    x = torch.from_numpy(exc).to(cuda_device, b.dtype)
    _, _, masks = b.step_many(x, 0.02, forces_every=1, return_masks=True)
    assert np.array_equal(masks.cpu().numpy(), want)
    b = MuscleBatch(spec, m, cuda_device, precision)
    mask = torch.empty(m, n, dtype=torch.uint8, device=cuda_device)
    for k in range(K):
        b.step(x[k], 0.02, mask_out=mask)
        assert np.array_equal(mask.cpu().numpy(), want[k])


@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_fused_equals_single_step_and_decimated_forces(cuda_device, precision):
    """K2 against K1 on the same inputs, odd sizes: M not a multiple of the block's muscle
    count, K not a multiple of the tile, windows of different lengths."""
    rng = np.random.default_rng(5)
    m, n = 203, 120
    spec = MuscleSpec.standard()
    exc = rng.random((97, m)).astype(np.float32)
    ptf = spec.tables()["ptf"]
    a = MuscleBatch(spec, m, cuda_device, precision)
    b = MuscleBatch(spec, m, cuda_device, precision)
    x = torch.from_numpy(exc).to(cuda_device, a.dtype)
    t1 = torch.stack([a.step(x[k], 0.02) for k in range(97)])
    f1_at = {}
    t2a, fdec = b.step_many(x[:60], 0.02, forces_every=7)
    t2b = b.step_many(x[60:], 0.02)
    close(torch.cat([t2a, t2b]).cpu().numpy(), t1.double().cpu().numpy(), precision, total_scale(spec), "totals")
    close(b.cpf.cpu().numpy(), a.cpf.cpu().numpy(), precision, ptf, "capacities")
    close(b.current_forces.cpu().numpy(), a.current_forces.double().cpu().numpy(), precision, ptf, "forces")
    assert fdec.shape == (60 // 7, m, n)
    c = MuscleBatch(spec, m, cuda_device, precision)
    for k in range(56):
        c.step(x[k], 0.02)
        if (k + 1) % 7 == 0:
            close(fdec[(k + 1) // 7 - 1].cpu().numpy(), c.current_forces.double().cpu().numpy(), precision, ptf, f"dec{k}")


@pytest.mark.parametrize("precision", ["f32", "f64"])
@pytest.mark.parametrize("upt", [1, 2, 4])
def test_fused_units_per_thread_variants(cuda_device, precision, upt):
    rng = np.random.default_rng(upt)
    m, n, K = 70, 120, 40
    exc = (67.0 * rng.random((K, m))).astype(np.float32)
    ora = OracleMuscles(n, m, "potvin")
    want = np.stack([ora.step(exc[k].astype(np.float64), 0.02).total for k in range(K)])
    b = MuscleBatch(MuscleSpec.potvin(n), m, cuda_device, precision)
    got = b.step_many(torch.from_numpy(exc).to(cuda_device, b.dtype), 0.02, units_per_thread=upt)
    close(got.cpu().numpy(), want, precision, 2609.03, "totals")
    close(b.cpf.cpu().numpy(), ora.cpf, precision, ora.t.ptf, "capacities")
    np.testing.assert_allclose(b.dur.cpu().numpy(), ora.dur, rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_fused_constant_excitation_and_unaligned_batch(cuda_device, precision):
    """exc_stride_k == 0 (one value per muscle for the window) and M % 4 != 0 (no bulk copies)."""
    rng = np.random.default_rng(8)
    for m in (64, 61):
        e = (67.0 * rng.random(m)).astype(np.float32)
        ora = OracleMuscles(120, m, "potvin")
        want = np.stack([ora.step(e.astype(np.float64), 0.02).total for _ in range(50)])
        b = MuscleBatch(MuscleSpec.potvin(120), m, cuda_device, precision)
        got = b.step_many(torch.from_numpy(e).to(cuda_device, b.dtype), 0.02, steps=50)
        close(got.cpu().numpy(), want, precision, 2609.03, f"const m={m}")
        b = MuscleBatch(MuscleSpec.potvin(120), m, cuda_device, precision)
        full = torch.from_numpy(np.repeat(e[None], 50, 0)).to(cuda_device, b.dtype)
        got = b.step_many(full, 0.02)
        close(got.cpu().numpy(), want, precision, 2609.03, f"per-step m={m}")


@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_halves_pool_and_fibers_stages(cuda_device, precision):
    """Pool.step and Fibers.step on their own (stage POOL / stage FIBERS of mb_step)."""
    g = golden("halves")
    n = 120
    b = MuscleBatch(MuscleSpec.potvin(n), 1, cuda_device, precision)
    rates = torch.empty(1, n, dtype=b.dtype, device=cuda_device)
    for e, want in zip(g["pool_exc"], g["pool_rates"]):
        b.step(torch.from_numpy(e.astype(np.float64)).to(cuda_device, b.dtype), 0.5, stage=nat.MB_STAGE_POOL, rates_out=rates)
        close(rates[0].cpu().numpy(), want, precision, 35.0, "rates")
    np.testing.assert_allclose(b.dur[0].cpu().numpy(), g["pool_dur"], rtol=1e-12)
    for name, spec in (("fib", MuscleSpec.potvin(n)), ("pmf", MuscleSpec(n, fibers=MuscleSpec.standard().fibers))):
        b = MuscleBatch(spec, 1, cuda_device, precision)
        ptf = spec.tables()["ptf"]
        for r, want in zip(g[f"{name}_rates"], g[f"{name}_totals"]):
            tot = b.step(torch.from_numpy(r.astype(np.float64)).to(cuda_device, b.dtype), 0.5, stage=nat.MB_STAGE_FIBERS)
            close(tot.cpu().numpy(), [want], precision, 2609.03, name + " total")
        close(b.cpf[0].cpu().numpy(), g[f"{name}_cpf"], precision, ptf, name + " cpf")
        close(b.current_forces[0].cpu().numpy(), g[f"{name}_forces_last"], precision, ptf, name + " forces")


@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_arm_ragged_rows(cuda_device, precision):
    """BASELINE config #3 shape: 30 StandardMuscles of 256..3127 units per row (35,032 units)."""
    g = golden("arm")
    specs = [MuscleSpec.standard(max_force=float(f)) for f in g["max_forces"]]
    assert [s.n for s in specs] == list(g["counts"])
    rows = 3
    b = MuscleBatch(specs, rows, cuda_device, precision)
    assert b.L == 35032 and b.S == 30
    ptf = np.concatenate([s.tables()["ptf"] for s in specs])
    for k in range(g["exc"].shape[0]):
        x = torch.from_numpy(np.repeat(g["exc"][k][None], rows, 0)).to(cuda_device, b.dtype)
        tot = b.step(x, float(g["dt"]))
        for r in range(rows):
            close(tot[r].cpu().numpy(), g["totals"][k], precision, 1.0, f"arm totals step {k}")
    for r in range(rows):
        close(b.cpf[r].cpu().numpy(), g["cpf"], precision, ptf, "arm capacities")
        close(b.current_forces[r].cpu().numpy(), g["forces"], precision, ptf, "arm forces")
    fat = b.get_peripheral_fatigue()
    want = np.array([1 - sum(g["cpf"][o:o + s.n]) / s.output_divisor
                     for o, s in zip(np.cumsum([0] + [s.n for s in specs[:-1]]), specs)])
    close(fat[0].cpu().numpy(), want, precision, 1.0, "arm peripheral fatigue")


def test_full_size_properties_config2(cuda_device):
    """BASELINE config #2 at full size (65,536 x 120, fused 1,000 steps, FP32 mode), checked
    through size-independent properties: (i) muscles given the same excitation produce the
    same output wherever they sit in the batch; (ii) a sample of muscles matches the oracle;
    (iii) capacities never increase for Potvin fibers and stay in [0, peak]; (iv) zero
    excitation leaves the state untouched."""
    M, n, K, dt = 65536, 120, 1000, 0.02
    rng = np.random.default_rng(2017)
    levels = (67.0 * rng.random(64)).astype(np.float32)
    levels[0] = 0.0
    idx = rng.integers(0, 64, size=M)
    idx[:64] = np.arange(64)
    e = torch.from_numpy(levels[idx]).to(cuda_device)
    b = MuscleBatch(MuscleSpec.potvin(n), M, cuda_device, "f32")
    tot = b.step_many(e, dt, steps=K)
    torch.cuda.synchronize()
    tot = tot.cpu().numpy()
    cpf = b.cpf.cpu().numpy()
    for lv in range(64):                                   # (i)
        cols = np.nonzero(idx == lv)[0]
        assert np.array_equal(tot[:, cols], np.repeat(tot[:, cols[:1]], len(cols), 1))
        assert np.array_equal(cpf[cols], np.repeat(cpf[cols[:1]], len(cols), 0))
    ora = OracleMuscles(n, 64, "potvin")                   # (ii)
    want = np.stack([ora.step(levels.astype(np.float64), dt).total for _ in range(K)])
    close(tot[:, :64], want, "f32", 2609.03, "sampled totals")
    close(cpf[:64], ora.cpf, "f32", ora.t.ptf, "sampled capacities")
    assert (cpf <= ora.t.ptf[None] * (1 + 1e-7)).all() and (cpf >= 0).all()      # (iii)
    zero = np.nonzero(idx == 0)[0]                          # (iv)
    assert np.array_equal(cpf[zero], np.repeat(ora.t.ptf[None], len(zero), 0))
    assert (tot[:, zero] == 0).all() and (b.dur.cpu().numpy()[zero] == 0).all()


def test_state_dict_roundtrip_and_fresh_outputs(cuda_device):
    spec = MuscleSpec.standard()
    a = MuscleBatch(spec, 16, cuda_device, "f64")
    x = torch.rand(16, device=cuda_device, dtype=torch.float64)
    f_prev = None
    for _ in range(10):
        a.step(x, 0.1)
        f = a.current_forces
        assert f_prev is None or f.data_ptr() != f_prev.data_ptr()      # fresh array per step (fibers:258)
        f_prev = f
    st = a.state_dict()
    b = MuscleBatch(spec, 16, cuda_device, "f64")
    b.load_state_dict(st)
    assert torch.equal(a.step(x, 0.1), b.step(x, 0.1))
    assert torch.equal(a.cpf, b.cpf)
    b.reset()
    assert torch.equal(b.cpf[0], torch.from_numpy(spec.tables()["ptf"]).to(cuda_device))


def test_full_size_properties_config5(cuda_device):
    """BASELINE config #5 at full size: StandardMuscle with 2,000 units x 131,072 instances, one
    launch per frame (the HBM streaming kernel), FP32 mode.  (i) equal inputs give equal
    outputs anywhere in the batch; (ii) a sample matches the oracle; (iii) capacities stay in
    [0, peak]; (iv) resting muscles keep their exact state."""
    M, dt, frames = 131072, 0.02, 4
    spec = MuscleSpec.standard(527.1)
    assert spec.n == 2000
    rng = np.random.default_rng(5)
    levels = rng.random((frames, 32)).astype(np.float32)
    levels[:, 0] = 0.0
    idx = rng.integers(0, 32, size=M)
    idx[:32] = np.arange(32)
    b = MuscleBatch(spec, M, cuda_device, "f32", fresh_outputs=False)
    ora = OracleMuscles.standard(527.1, m=32)
    ptf = ora.t.ptf
    for k in range(frames):
        tot = b.step(torch.from_numpy(levels[k][idx]).to(cuda_device), dt)
        want = ora.step(levels[k].astype(np.float64), dt)
        tot = tot.cpu().numpy()
        for lv in (0, 7, 31):                                                  # (i)
            cols = np.nonzero(idx == lv)[0]
            assert (tot[cols] == tot[cols[0]]).all()
        close(tot[:32], want.total, "f32", 1.0, f"sampled totals, frame {k}")  # (ii)
        close(b.current_forces[:32].cpu().numpy(), want.forces, "f32", ptf, f"sampled forces, frame {k}")
    cpf = b.cpf
    close(cpf[:32].cpu().numpy(), ora.cpf, "f32", ptf, "sampled capacities")
    peak = torch.from_numpy(ptf).to(cuda_device)
    assert bool((cpf <= peak[None]).all()) and bool((cpf >= 0).all())         # (iii)
    zero = torch.from_numpy(np.nonzero(idx == 0)[0]).to(cuda_device)            # (iv)
    assert bool((cpf[zero] == peak[None]).all())


def test_full_size_properties_config3(cuda_device):
    """BASELINE config #3 at full size: the 30-muscle arm (35,032 units) x 4,096 agents, per-step
    random excitation per (agent, muscle), single-step launches on the ragged layout, FP32 mode."""
    g = golden("arm")
    specs = [MuscleSpec.standard(max_force=float(f)) for f in g["max_forces"]]
    agents, dt, frames = 4096, 0.02, 3
    rng = np.random.default_rng(3)
    b = MuscleBatch(specs, agents, cuda_device, "f32", fresh_outputs=False)
    assert b.n_units == 4096 * 35032 and b.n_muscles == 122880
    oras = [OracleMuscles.standard(float(f), m=4) for f in g["max_forces"]]
    for k in range(frames):
        exc = rng.random((agents, 30)).astype(np.float32)
        exc[1] = exc[0]                                                         # two identical agents
        tot = b.step(torch.from_numpy(exc).to(cuda_device), dt).cpu().numpy()
        assert tot.shape == (agents, 30) and np.array_equal(tot[0], tot[1])
        for s, o in enumerate(oras):                                            # agents 0..3 against the oracle
            want = o.step(exc[:4, s].astype(np.float64), dt)
            close(tot[:4, s], want.total, "f32", 1.0, f"arm totals, muscle {s}, frame {k}")
    off = np.cumsum([0] + [s.n for s in specs])
    cpf = b.cpf[:4].cpu().numpy()
    for s, o in enumerate(oras):
        close(cpf[:, off[s]:off[s + 1]], o.cpf, "f32", o.t.ptf, f"arm capacities, muscle {s}")
    assert np.array_equal(b.cpf[0].cpu().numpy(), b.cpf[1].cpu().numpy())


def test_full_size_properties_config4_shard(cuda_device):
    """BASELINE config #4, one GPU's shard (32,768 PotvinFuglevandMuscle(120), FP64 mode): the
    ramp-hold-rest protocol compressed to 1,500 steps, run as fused windows of 500; a sample of
    muscles against the oracle at rtol 1e-9, masks of the last step bit-exact."""
    M, n, dt, K = 32768, 120, 0.02, 1500
    rng = np.random.default_rng(4)
    e_max = 67.0 * (0.5 + 0.5 * rng.random(M))
    k = np.arange(K)
    shape = np.where(k < 300, k / 300.0, np.where(k < 900, 1.0, 0.0))          # ramp, hold, rest
    b = MuscleBatch(MuscleSpec.potvin(n), M, cuda_device, "f64", fresh_outputs=False)
    ora = OracleMuscles(n, 16, "potvin")
    em = torch.from_numpy(e_max).to(cuda_device)
    sh = torch.from_numpy(shape).to(cuda_device)
    tots = []
    for w in range(0, K, 500):
        exc = sh[w:w + 500, None] * em[None, :]
        tots.append(b.step_many(exc, dt)[:, :16].cpu().numpy())
    want = np.stack([ora.step(shape[i] * e_max[:16], dt).total for i in range(K)])
    close(np.concatenate(tots), want, "f64", 2609.03, "sampled totals")
    close(b.cpf[:16].cpu().numpy(), ora.cpf, "f64", ora.t.ptf, "sampled capacities")
    np.testing.assert_allclose(b.dur[:16].cpu().numpy(), ora.dur, rtol=1e-9, atol=1e-12)
    assert bool((b.cpf <= torch.from_numpy(ora.t.ptf).to(cuda_device)[None]).all())


@pytest.mark.parametrize("overlap", [False, True])
def test_host_pipeline_matches_the_device_path(cuda_device, overlap):
    """step_many_host (pinned host buffers in / out, H2D / kernel / D2H pipelined over sub-windows) is
    what bench.py's end-to-end leg times: it must return exactly what step_many returns, window
    after window, with and without cross-call overlap, for chunk sizes that do not divide K."""
    M, n, K, dt = 300, 120, 200, 0.02
    rng = np.random.default_rng(8)
    a = MuscleBatch(MuscleSpec.potvin(n), M, cuda_device, "f32", fresh_outputs=False)
    b = MuscleBatch(MuscleSpec.potvin(n), M, cuda_device, "f32", fresh_outputs=False)
    out_h = [torch.empty(K, M).pin_memory() for _ in range(3)]
    excs = [torch.from_numpy((67.0 * rng.random((K, M))).astype(np.float32)).pin_memory() for _ in range(3)]
    for w in range(3):
        a.step_many_host(excs[w], dt, out_h[w], chunk=72, overlap_calls=overlap)
    a.host_pipeline_join()
    torch.cuda.synchronize()
    for w in range(3):
        # same sub-window partition on the device path: bitwise equal (FP32-mode rounding depends on
        # where a launch ends, because the state goes back to float64 there)
        x = excs[w].to(cuda_device)
        want = torch.cat([b.step_many(x[k0:k0 + 72], dt) for k0 in range(0, K, 72)])
        assert torch.equal(out_h[w], want.cpu()), f"window {w}"
    assert torch.equal(a.cpf, b.cpf) and torch.equal(a.dur, b.dur)
    # and against one undivided launch: within the FP32-mode bar
    c = MuscleBatch(MuscleSpec.potvin(n), M, cuda_device, "f32", fresh_outputs=False)
    whole = c.step_many(excs[0].to(cuda_device), dt)
    close(out_h[0].numpy(), whole.double().cpu().numpy(), "f32", 2609.03, "sub-windows vs one window")


@pytest.mark.parametrize("precision", ["f32", "f64"])
@pytest.mark.parametrize("kind,n", [("potvin", 121), ("potvin", 57), ("standard", 299), ("standard", 1999)])
def test_odd_row_lengths_in_the_streaming_kernel(cuda_device, kind, n, precision):
    """Odd unit counts: consecutive float64 state rows alternate between 16-byte aligned and not, so the
    streaming kernel's bulk copies start one element early in every other row.  Odd number of rows too
    (partial last chunk, odd element count overall), per-muscle and per-unit inputs, fatigue epilogue."""
    from pymuscle_b200.tables import fibers_tables
    rng = np.random.default_rng(n)
    rows, K = 11, 25
    if kind == "potvin":
        spec, scale = MuscleSpec.potvin(n), 67.0
    else:
        base = MuscleSpec.standard()
        spec, scale = MuscleSpec(n, base.pool, base.fibers, input_scale=base.input_scale,
                                 output_divisor=sum(fibers_tables(n, base.fibers)[0]), name="standard"), 1.0
    ora, ora_u = OracleMuscles(n, rows, kind), OracleMuscles(n, rows, kind)
    ptf = spec.tables()["ptf"]
    tscale = float(np.sum(ptf)) / spec.output_divisor
    b, bu = MuscleBatch(spec, rows, cuda_device, precision), MuscleBatch(spec, rows, cuda_device, precision)
    for k in range(K):
        e = (scale * rng.random(rows)).astype(np.float32)
        eu = (scale * rng.random((rows, n))).astype(np.float32)
        want, want_u = ora.step(e.astype(np.float64), 0.05), ora_u.step(eu.astype(np.float64), 0.05)
        if k % 2:
            got, fat = b.step_env(torch.from_numpy(e).to(cuda_device, b.dtype), 0.05)       # fused fatigue epilogue
            if kind == "standard":
                close(fat.cpu().numpy(), ora.peripheral_fatigue(), precision, 1.0, f"fatigue step {k}")
        else:
            got = b.step(torch.from_numpy(e).to(cuda_device, b.dtype), 0.05)
        got_u = bu.step(torch.from_numpy(eu).to(cuda_device, b.dtype), 0.05)
        close(got.cpu().numpy(), want.total, precision, tscale, f"totals step {k}")
        close(got_u.cpu().numpy(), want_u.total, precision, tscale, f"per-unit-input totals step {k}")
        close(b.current_forces.cpu().numpy(), want.forces, precision, ptf, f"forces step {k}")
        close(bu.current_forces.cpu().numpy(), want_u.forces, precision, ptf, f"per-unit-input forces step {k}")
    close(b.cpf.cpu().numpy(), ora.cpf, precision, ptf, "capacities")
    close(bu.cpf.cpu().numpy(), ora_u.cpf, precision, ptf, "per-unit-input capacities")
    np.testing.assert_allclose(b.dur.cpu().numpy(), ora.dur, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(bu.dur.cpu().numpy(), ora_u.dur, rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_ragged_rows_with_an_odd_row_length(cuda_device, precision):
    """Three StandardMuscles of 256 + 280 + 305 = 841 units per row: odd row length AND odd segment
    offsets, five rows (partial chunk); totals, forces, capacities and the fused fatigue epilogue."""
    forces = [round(68.0 * (824.0 / 68.0) ** (i / 29.0), 1) for i in range(3)]
    specs = [MuscleSpec.standard(f) for f in forces]
    assert sum(s.n for s in specs) == 841
    rows, rng = 5, np.random.default_rng(3)
    oras = [OracleMuscles.standard(f, m=rows) for f in forces]
    b = MuscleBatch(specs, rows, cuda_device, precision)
    offs = np.cumsum([0] + [s.n for s in specs])
    for k in range(20):
        e = rng.random((rows, 3)).astype(np.float32)
        got, fat = b.step_env(torch.from_numpy(e).to(cuda_device, b.dtype), 0.1)
        for j, (o, s) in enumerate(zip(oras, specs)):
            want = o.step(e[:, j].astype(np.float64), 0.1)
            ptf = s.tables()["ptf"]
            close(got[:, j].cpu().numpy(), want.total, precision, 1.0, f"totals muscle {j} step {k}")
            close(b.current_forces[:, offs[j]:offs[j + 1]].cpu().numpy(), want.forces, precision, ptf, f"forces {j}/{k}")
            close(b.cpf[:, offs[j]:offs[j + 1]].cpu().numpy(), o.cpf, precision, ptf, f"capacities {j}/{k}")
            close(fat[:, j].cpu().numpy(), o.peripheral_fatigue(), precision, 1.0, f"fatigue {j}/{k}")


@pytest.mark.parametrize("precision", ["f32", "f64"])
@pytest.mark.parametrize("kind", ["potvin", "standard"])
def test_generic_single_step_kernel_agrees_with_the_streaming_kernel(cuda_device, kind, precision, monkeypatch):
    """The generic k1_step kernel is what mask / rate outputs, stage launches and host-resident inputs use;
    for the whole-muscle stage it must give what the streaming kernel gives (MB_K1_GENERIC=1 forces it).
    State and forces are per-unit results of the same arithmetic: bitwise equal; totals differ only by the
    order of the per-muscle sum."""
    rng = np.random.default_rng(17)
    rows, n, K = 19, 120, 12
    spec = MuscleSpec.potvin(n) if kind == "potvin" else MuscleSpec.standard()
    scale = 67.0 if kind == "potvin" else 1.0
    a, b = MuscleBatch(spec, rows, cuda_device, precision), MuscleBatch(spec, rows, cuda_device, precision)
    for k in range(K):
        x = torch.from_numpy((scale * rng.random(rows)).astype(np.float32)).to(cuda_device, a.dtype)
        ta = a.step(x, 0.05)
        monkeypatch.setenv("MB_K1_GENERIC", "1")
        tb = b.step(x, 0.05)
        monkeypatch.delenv("MB_K1_GENERIC")
        if precision == "f32":      # FP64 mode: the streaming kernel uses the table-based exp, the generic one the library's
            assert torch.equal(a.cpf, b.cpf) and torch.equal(a.dur, b.dur)
            assert torch.equal(a.current_forces, b.current_forces)
        close(tb.cpu().numpy(), ta.double().cpu().numpy(), precision, total_scale(spec), f"totals step {k}")
    close(b.cpf.cpu().numpy(), a.cpf.cpu().numpy(), precision, spec.tables()["ptf"], "capacities")
    close(b.current_forces.cpu().numpy(), a.current_forces.double().cpu().numpy(), precision, spec.tables()["ptf"], "forces")


@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_many_identical_rows_give_identical_results(cuda_device, precision):
    """The streaming kernel's work-item cursor (blocks stride through (chunk, segment) pairs without a
    division) and the per-row alignment shifts, at a size where every block walks many items: all rows
    get the same excitations, so every row must reproduce row 0 bit for bit, and row 0 the oracle.
    Uniform odd layout (4,099 rows x 1,999 units) and the ragged arm (1,003 rows x 30 muscles)."""
    rng = np.random.default_rng(23)
    # uniform, odd row length, odd row count
    from pymuscle_b200.tables import fibers_tables
    base = MuscleSpec.standard()
    n, rows = 1999, 4099
    spec = MuscleSpec(n, base.pool, base.fibers, input_scale=base.input_scale,
                      output_divisor=sum(fibers_tables(n, base.fibers)[0]), name="standard")
    b, ora = MuscleBatch(spec, rows, cuda_device, precision), OracleMuscles(n, 1, "standard")
    for k in range(4):
        e = float(np.float32(rng.random()))                   # representable in both modes
        tot = b.step(torch.full((rows,), e, device=cuda_device, dtype=b.dtype), 0.1)
        want = ora.step(np.array([e]), 0.1)
        assert torch.equal(tot, tot[:1].expand(rows))
        assert torch.equal(b.cpf, b.cpf[:1].expand(rows, n)) and torch.equal(b.current_forces, b.current_forces[:1].expand(rows, n))
        close(tot[:1].cpu().numpy(), want.total, precision, 1.0, f"uniform totals step {k}")
    close(b.cpf[0].cpu().numpy(), ora.cpf[0], precision, spec.tables()["ptf"], "uniform capacities")
    del b
    # ragged arm
    g = golden("arm")
    specs = [MuscleSpec.standard(max_force=float(f)) for f in g["max_forces"]]
    rows = 1003
    b = MuscleBatch(specs, rows, cuda_device, precision)
    ptf = np.concatenate([s.tables()["ptf"] for s in specs])
    for k in range(g["exc"].shape[0]):
        x = torch.from_numpy(np.repeat(g["exc"][k][None], rows, 0)).to(cuda_device, b.dtype)
        tot, fat = b.step_env(x, float(g["dt"]))
        assert torch.equal(tot, tot[:1].expand(rows, 30)) and torch.equal(fat, fat[:1].expand(rows, 30))
        close(tot[0].cpu().numpy(), g["totals"][k], precision, 1.0, f"arm totals step {k}")
    assert torch.equal(b.cpf, b.cpf[:1].expand(rows, b.L))
    close(b.cpf[0].cpu().numpy(), g["cpf"], precision, ptf, "arm capacities")
