"""
GPU tests added in round 2 (all through the C ABI):

* the gather of totals fused into the kernels' epilogue (``mb_step_fused_gather``) with the peer
  buffers pointing at LOCAL memory, so that a single GPU exercises every peer-store branch;
* held excitations (``exc_hold``) and decimated totals (``totals_every``) of ``mb_step_fused_ex``;
* the host-buffer pipeline's constant / held forms;
* a pool whose peak firing rates drop below ``min_firing_rate - derecruitment_delta`` (the fused fast
  path must not be taken: the reference clamps the then-negative adaptation, pool:205-206);
* the Hill-type curves evaluated inside the step kernel (hill_type.py:15-66);
* the fused fatigue readout at rest (exactly 0, muscle.py:248-256) and against the stand-alone kernel;
* history capture through the pinned-host ring (README.md:128-145);
* guard bands of NaN around the state tensors (no kernel writes outside its arrays).
"""
import os
import subprocess
import sys
from dataclasses import replace

import numpy as np
import pytest
import torch

from conftest import ROOT, golden
from oracle.pymuscle_oracle import OracleMuscles, PoolHyper
from pymuscle_b200 import ForceHistory, MuscleBatch, MuscleGroupEnv, MuscleSpec
from pymuscle_b200 import hill_type
from test_gpu_parity import close, total_scale

pytestmark = pytest.mark.gpu


def oracle_window(ora, exc, dt):
    return np.stack([ora.step(exc[k].astype(np.float64), dt).total for k in range(exc.shape[0])])


# --------------------------------------------------------------------------------------------------
# fused gather, single GPU: peers are local buffers
# --------------------------------------------------------------------------------------------------
def _batch_for(kernel, M, dev):
    """(batch, oracle factory, precision) routed to the named fused kernel."""
    if kernel == "k2w_f32":                       # one warp per muscle pair
        return MuscleBatch(MuscleSpec.potvin(120), M, dev, "f32"), lambda: OracleMuscles(120, M, "potvin"), "f32"
    if kernel == "k2_fused_f32":                  # block kernel: more than 128 units
        spec = MuscleSpec.standard(90.0)
        assert spec.n == 340
        return MuscleBatch(spec, M, dev, "f32"), lambda: OracleMuscles.standard(90.0, m=M), "f32"
    if kernel == "k2_fused_f64":
        return MuscleBatch(MuscleSpec.potvin(120), M, dev, "f64"), lambda: OracleMuscles(120, M, "potvin"), "f64"
    raise AssertionError(kernel)


@pytest.mark.parametrize("layout", ["step_major", "pair_major", "muscle_major"])
@pytest.mark.parametrize("n_peers", [1, 3])
@pytest.mark.parametrize("kernel", ["k2w_f32", "k2_fused_f32", "k2_fused_f64", "k2u_f32"])
def test_fused_gather_into_local_peer_buffers(cuda_device, kernel, n_peers, layout):
    M, K, dt = 37, 45, 0.02                       # odd muscle count: the last warp / pair is half empty
    rng = np.random.default_rng(5)
    if kernel == "k2u_f32":                       # per-instance tables: one warp per muscle
        base = MuscleSpec.potvin(120)
        specs = [replace(base, pool=replace(base.pool, adaptation_magnitude=0.5 + 0.01 * r)) for r in range(M)]
        b = MuscleBatch(base, M, cuda_device, "f32", row_specs=specs)
        oras = [OracleMuscles(120, 1, "potvin", pool=PoolHyper(adaptation_magnitude=0.5 + 0.01 * r)) for r in range(M)]
        precision, scale_in = "f32", 67.0
    else:
        b, mk, precision = _batch_for(kernel, M, cuda_device)
        oras, scale_in = None, (1.0 if kernel == "k2_fused_f32" else 67.0)
        ora = mk()
    exc = (scale_in * rng.random((K, M))).astype(np.float32)
    x = torch.from_numpy(exc).to(cuda_device, b.dtype)
    local = torch.empty(K, M, dtype=b.dtype, device=cuda_device)
    pairs, M_total, col0 = (M + 1) // 2, 2 * M + 6, 4            # this "rank" owns columns 4 .. 4+M of a wider gather
    nan = float("nan")
    if layout == "step_major":                                     # [K, M_total]: (stride_k, stride_pair) = (M_total, 2)
        bufs = [torch.full((K, M_total), nan, dtype=b.dtype, device=cuda_device) for _ in range(n_peers)]
        ptrs = [t.data_ptr() + col0 * t.element_size() for t in bufs]
        sk, sp = M_total, 2
    elif layout == "pair_major":                                   # region[pair][k][2]: (2, 2 * K_max, 1)
        kmax = K + 3
        bufs = [torch.full((pairs, kmax, 2), nan, dtype=b.dtype, device=cuda_device) for _ in range(n_peers)]
        ptrs = [t.data_ptr() for t in bufs]
        sk, sp, so = 2, 2 * kmax, 1
    else:                                                          # region[muscle][k]: (1, 2 * K_max, K_max)
        kmax = K + 3
        bufs = [torch.full((2 * pairs, kmax), nan, dtype=b.dtype, device=cuda_device) for _ in range(n_peers)]
        ptrs = [t.data_ptr() for t in bufs]
        sk, sp, so = 1, 2 * kmax, kmax
    if layout == "step_major":
        so = 1
    b.step_many_gather(x, dt, local, ptrs, sk, sp, so)
    torch.cuda.synchronize()
    got = local.cpu().numpy()
    if oras is None:
        want = oracle_window(ora, exc, dt)
        spec = b.specs[0]
    else:
        want = np.concatenate([oracle_window(o, exc[:, r:r + 1], dt) for r, o in enumerate(oras)], axis=1)
        spec = b.specs[0]
    close(got, want, precision, total_scale(spec), f"{kernel} local totals")
    for t in bufs:
        h = t.cpu().numpy()
        if layout == "step_major":
            assert np.array_equal(h[:, col0:col0 + M], got), "peer copy differs from the local totals"
            assert np.isnan(h[:, :col0]).all() and np.isnan(h[:, col0 + M:]).all(), "stores outside this rank's region"
        elif layout == "pair_major":
            dense = h[:, :K].transpose(1, 0, 2).reshape(K, -1)
            assert np.array_equal(dense[:, :M], got)
            assert np.isnan(h[:, K:]).all(), "stores beyond the window"
            if M & 1:
                assert np.isnan(dense[:, M:]).all(), "the empty half of the last pair was written"
        else:
            assert np.array_equal(h[:M, :K].T, got)
            assert np.isnan(h[:, K:]).all() and np.isnan(h[M:]).all(), "stores outside the muscles' rows"


# --------------------------------------------------------------------------------------------------
# held excitations / decimated totals
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision,kind,n", [("f32", "potvin", 120), ("f64", "potvin", 120), ("f32", "standard", 340),
                                              ("f32", "potvin", 33)])
@pytest.mark.parametrize("hold,tev", [(1, 4), (5, 1), (3, 6), (7, 5)])
def test_hold_and_totals_every_match_the_plain_window(cuda_device, precision, kind, n, hold, tev):
    M, K, dt = 19, 90, 0.02
    spec = MuscleSpec.potvin(n) if kind == "potvin" else MuscleSpec.standard(90.0)
    assert spec.n == n
    rng = np.random.default_rng(hold * 10 + tev)
    rows = (K + hold - 1) // hold
    ctrl = ((67.0 if kind == "potvin" else 1.0) * rng.random((rows, M))).astype(np.float32)
    a = MuscleBatch(spec, M, cuda_device, precision)
    bref = MuscleBatch(spec, M, cuda_device, precision)
    xc = torch.from_numpy(ctrl).to(cuda_device, a.dtype)
    got = a.step_many(xc, dt, steps=K, hold=hold, totals_every=tev)
    full = bref.step_many(xc.repeat_interleave(hold, dim=0)[:K].contiguous(), dt)
    torch.cuda.synchronize()
    assert tuple(got.shape) == (K // tev, M)
    assert torch.equal(got, full[tev - 1::tev][:K // tev]), "decimated / held window differs from the plain one"
    assert torch.equal(a.cpf, bref.cpf) and torch.equal(a.dur, bref.dur)
    assert torch.equal(a.current_forces, bref.current_forces)


def test_ragged_fallback_honours_hold_and_totals_every(cuda_device):
    specs = [MuscleSpec.standard(f) for f in (20.0, 32.0, 45.0)]
    rows, K, dt, hold, tev = 5, 24, 0.02, 4, 3
    a = MuscleBatch(specs, rows, cuda_device, "f64")
    ref = MuscleBatch(specs, rows, cuda_device, "f64")
    ctrl = torch.rand(K // hold, rows * 3, device=cuda_device, dtype=torch.float64)
    got = a.step_many(ctrl, dt, steps=K, hold=hold, totals_every=tev)
    want = torch.stack([ref.step(ctrl[k // hold].view(rows, 3), dt).reshape(-1) for k in range(K)])
    torch.testing.assert_close(got, want[tev - 1::tev], rtol=1e-9, atol=1e-12)   # (one launch vs K launches: sums may re-associate)
    torch.testing.assert_close(a.cpf, ref.cpf, rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("form", ["per_step", "constant", "held"])
def test_host_pipeline_forms(cuda_device, form):
    M, K, dt = 256, 70, 0.02
    spec = MuscleSpec.potvin(120)
    a, ref = MuscleBatch(spec, M, cuda_device, "f32"), MuscleBatch(spec, M, cuda_device, "f32")
    g = torch.Generator().manual_seed(3)
    hold = 5 if form == "held" else 1
    tev = 2 if form != "per_step" else 1
    rows = 1 if form == "constant" else (K + hold - 1) // hold
    ctrl = (67.0 * torch.rand(rows, M, generator=g)).pin_memory()
    out = torch.full((K // tev, M), float("nan")).pin_memory()
    for _ in range(2):                                               # two windows: the pipeline carries over calls
        if form == "constant":
            a.step_many_host(ctrl[0], dt, out, chunk=32, steps=K, totals_every=tev, overlap_calls=True)
            want = ref.step_many(ctrl[0].to(cuda_device), dt, steps=K)
        elif form == "held":
            a.step_many_host(ctrl, dt, out, chunk=32, steps=K, hold=hold, totals_every=tev, overlap_calls=True)
            want = ref.step_many(ctrl.to(cuda_device).repeat_interleave(hold, dim=0)[:K].contiguous(), dt)
        else:
            a.step_many_host(ctrl, dt, out, chunk=32, overlap_calls=True)
            want = ref.step_many(ctrl.to(cuda_device), dt)
        a.host_pipeline_join()
        torch.cuda.synchronize()
        # sub-windows re-enter the kernel through the float64 state (the float pair and the adaptation argument are
        # re-derived from it), so chunked and unchunked windows agree to rounding, not bit for bit
        torch.testing.assert_close(out, want[tev - 1::tev][:K // tev].cpu(), rtol=2e-5, atol=1e-4)
    torch.testing.assert_close(a.cpf, ref.cpf, rtol=1e-6, atol=1e-9)


# --------------------------------------------------------------------------------------------------
# peak rates below minR - d  =>  negative adaptation must be clamped (pool:205-206)
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_low_peak_rates_take_the_general_adaptation_path(cuda_device, precision):
    n, M, K, dt = 120, 8, 400, 0.02
    base = MuscleSpec.potvin(n)
    spec = replace(base, pool=replace(base.pool, max_firing_rate_last_unit=5))
    ph = PoolHyper(max_firing_rate_last_unit=5)
    ora = OracleMuscles(n, M, "potvin", pool=ph)
    assert (ora.t.peak < ph.min_firing_rate - ph.derecruitment_delta).any(), "the case needs units whose peak is below minR - d"
    rng = np.random.default_rng(11)
    exc = (spec.pool.max_excitation * (0.6 + 0.4 * rng.random((K, M)))).astype(np.float32)
    b = MuscleBatch(spec, M, cuda_device, precision)
    from pymuscle_b200 import _native as nat
    assert not (b.cfg.table_props & nat.MB_PROP_ADAPT_NONNEG)
    got = b.step_many(torch.from_numpy(exc).to(cuda_device, b.dtype), dt)
    want = oracle_window(ora, exc, dt)
    close(got.cpu().numpy(), want, precision, float(np.sum(ora.t.ptf)), "totals with clamped adaptation")
    close(b.cpf.cpu().numpy(), ora.cpf, precision, ora.t.ptf, "capacities")
    close(b.current_forces.cpu().numpy(), ora.current_forces, precision, ora.t.ptf, "forces")


# --------------------------------------------------------------------------------------------------
# Hill-type curves inside the step kernel
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_hill_epilogue_matches_the_reference_curves(cuda_device, precision):
    specs = [MuscleSpec.standard(f) for f in (20.0, 32.0, 45.0, 32.0)]
    agents, dt = 33, 0.02
    rng = np.random.default_rng(2)
    rest = rng.uniform(0.8, 1.2, (agents, 4))
    cur = rest * rng.uniform(0.7, 1.5, (agents, 4))
    prev = cur * rng.uniform(0.97, 1.03, (agents, 4))
    act = rng.random((agents, 4))
    a = MuscleBatch(specs, agents, cuda_device, precision)
    ref = MuscleBatch(specs, agents, cuda_device, precision)
    t = lambda v: torch.from_numpy(v).to(cuda_device, a.dtype)
    out, fat = a.step_env(t(act), dt, hill=(t(rest), t(cur), t(prev)))
    plain, fat0 = ref.step_env(t(act), dt)
    # the reference's own functions on float64 NumPy inputs (bit-identical to pymuscle/hill_type.py, tests/golden/hill.npz)
    gain = hill_type.contractile_element_force_length_curve(rest, cur) * \
        hill_type.contractile_element_force_velocity_curve(rest, cur, prev, dt)
    want = plain.double().cpu().numpy() * gain
    rtol = 1e-9 if precision == "f64" else 2e-5
    np.testing.assert_allclose(out.double().cpu().numpy(), want, rtol=rtol, atol=1e-300)
    assert torch.equal(fat, fat0) and torch.equal(a.cpf, ref.cpf)


def test_hill_golden_values_through_the_kernel(cuda_device):
    """The curve samples of tests/golden/hill.npz (produced by the unmodified reference), one muscle per sample."""
    g = golden("hill")
    rest, cur, prev, dt = g["rest"], g["cur"], g["prev"], float(g["dt"])
    m = rest.size
    b = MuscleBatch(MuscleSpec.standard(), m, cuda_device, "f64")
    ref = MuscleBatch(MuscleSpec.standard(), m, cuda_device, "f64")
    t = lambda v: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64)).to(cuda_device).view(m, 1)
    x = torch.full((m,), 0.7, dtype=torch.float64, device=cuda_device)
    out, _ = b.step_env(x, dt, hill=(t(rest), t(cur), t(prev)))
    plain, _ = ref.step_env(x, dt)
    # atol: 1.8 - 1.8/(1 + e) cancels near 0 (see tests/test_hill_type.py)
    np.testing.assert_allclose(out.cpu().numpy(), plain.cpu().numpy() * g["fl"] * g["fv"], rtol=1e-9, atol=1e-14)


def test_env_adapter_uses_the_fused_hill_epilogue(cuda_device):
    specs = [MuscleSpec.standard(f) for f in (20.0, 32.0)]
    agents, dt = 16, 0.02
    env = MuscleGroupEnv(specs, agents, cuda_device, "f64", rest_lengths=torch.tensor([1.0, 0.9], dtype=torch.float64))
    ora = [OracleMuscles.standard(20.0, m=agents), OracleMuscles.standard(32.0, m=agents)]
    rng = np.random.default_rng(8)
    rest = np.array([1.0, 0.9])
    prev = None
    launches0 = env.batch.launches
    for step in range(6):
        act = rng.random((agents, 2))
        cur = rest * rng.uniform(0.85, 1.3, (agents, 2))
        out, fat = env.step(torch.from_numpy(act), dt, lengths=torch.from_numpy(cur))
        p = cur if prev is None else prev
        gain = hill_type.contractile_element_force_length_curve(rest, cur) * \
            hill_type.contractile_element_force_velocity_curve(rest, cur, p, dt)
        want = np.stack([o.step(act[:, i], dt).total for i, o in enumerate(ora)], axis=1) * gain
        np.testing.assert_allclose(out.cpu().numpy(), want, rtol=1e-9, atol=1e-14)
        np.testing.assert_allclose(fat.cpu().numpy(), np.stack([o.peripheral_fatigue() for o in ora], axis=1),
                                   rtol=1e-9, atol=1e-14)
        prev = cur
    assert env.batch.launches - launches0 == 6, "an env step must be exactly one launch"


# --------------------------------------------------------------------------------------------------
# fused fatigue readout
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_fused_fatigue_is_exactly_zero_at_rest_and_equals_the_standalone_kernel(cuda_device, precision):
    specs = [MuscleSpec.standard(f) for f in (20.0, 32.0, 45.0)]
    rows, dt = 21, 0.02
    b = MuscleBatch(specs, rows, cuda_device, precision)
    zero = torch.zeros(rows, 3, dtype=b.dtype, device=cuda_device)
    _, fat = b.step_env(zero, dt)
    assert torch.count_nonzero(fat) == 0, "a rested muscle must report fatigue exactly 0 (muscle.py:248-256)"
    g = torch.Generator(device=cuda_device).manual_seed(1)
    for _ in range(30):
        _, fat = b.step_env(torch.rand(rows, 3, device=cuda_device, generator=g, dtype=b.dtype), dt)
    alone = b.get_peripheral_fatigue()
    # (the fused readout sums the lost capacity in the mode's arithmetic, the stand-alone kernel in float64)
    tol = dict(rtol=1e-12, atol=1e-15) if precision == "f64" else dict(rtol=2e-6, atol=1e-9)
    torch.testing.assert_close(fat, alone, **tol)
    off = np.concatenate([[0], np.cumsum([sp.n for sp in specs])])
    cpf = b.cpf.cpu().numpy()
    ref = np.stack([1.0 - np.array([sum(cpf[r, off[i]:off[i + 1]]) for r in range(rows)]) / specs[i].output_divisor
                    for i in range(3)], axis=1)                       # muscle.py:248-256, built-in sequential sum
    np.testing.assert_allclose(alone.cpu().numpy(), ref, rtol=1e-9, atol=1e-14)
    assert (fat > 0).all()


# --------------------------------------------------------------------------------------------------
# history capture
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_force_history_ring_matches_per_step_reads(cuda_device, precision):
    spec, M, dt, every, window = MuscleSpec.standard(), 3, 0.02, 5, 40
    b = MuscleBatch(spec, M, cuda_device, precision)
    ref = MuscleBatch(spec, M, cuda_device, precision)
    hist = ForceHistory(b, every=every, window=window, slots=2)
    x = torch.tensor([0.6, 0.3, 0.9], dtype=b.dtype, device=cuda_device)
    totals = []
    for _ in range(5):                                                # 5 windows through a 2-slot ring: it wraps
        totals.append(hist.advance(x, dt, steps=window))
    rec = hist.collect()
    want, want_tot = [], []
    for k in range(5 * window):
        want_tot.append(ref.step(x, dt))
        if (k + 1) % every == 0:
            want.append(ref.current_forces.double().cpu().numpy())
    want = np.stack(want)
    assert rec.shape == (5 * window // every, M, spec.n)
    tol = dict(rtol=1e-9, atol=1e-12) if precision == "f64" else dict(rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(rec, want, **tol)
    np.testing.assert_allclose(torch.cat(totals).double().cpu().numpy(), torch.stack(want_tot).double().cpu().numpy(), **tol)
    assert hist.collect().shape[0] == 0


# --------------------------------------------------------------------------------------------------
# memory safety without the sanitizer: NaN guard bands around the state
# --------------------------------------------------------------------------------------------------
def _guarded(batch, pad=64):
    """Move cpf / dur / forces of a batch into the middle of NaN-filled arenas (16-byte aligned views)."""
    arenas = {}
    for name in ("cpf", "dur"):
        t = getattr(batch, name)
        arena = torch.full((t.numel() + 2 * pad,), float("nan"), dtype=t.dtype, device=t.device)
        view = arena[pad:pad + t.numel()].view_as(t)
        view.copy_(t)
        setattr(batch, name, view)
        arenas[name] = arena
    return arenas


def _bands_intact(arenas, pad=64):
    for name, a in arenas.items():
        assert torch.isnan(a[:pad]).all() and torch.isnan(a[-pad:]).all(), f"{name}: a kernel wrote outside the array"


@pytest.mark.parametrize("precision", ["f32", "f64"])
@pytest.mark.parametrize("shape", ["odd_uniform", "ragged_odd", "tiny", "wide"])
def test_kernels_do_not_write_outside_their_state_arrays(cuda_device, precision, shape):
    dt = 0.02
    if shape == "odd_uniform":
        spec, rows = MuscleSpec.potvin(33), 4099                      # odd row length, partial last chunk
    elif shape == "ragged_odd":
        spec, rows = [MuscleSpec.standard(f) for f in (9.0, 33.0, 21.0)], 2051
    elif shape == "tiny":
        spec, rows = MuscleSpec.potvin(2), 3
    else:
        spec, rows = MuscleSpec.standard(527.1), 259
    b = MuscleBatch(spec, rows, cuda_device, precision)
    arenas = _guarded(b)
    ref = MuscleBatch(spec, rows, cuda_device, precision)
    g = torch.Generator(device=cuda_device).manual_seed(0)
    scale = 67.0 if shape in ("odd_uniform", "tiny") else 1.0
    for _ in range(3):
        x = scale * torch.rand(rows, b.S, device=cuda_device, generator=g, dtype=b.dtype)
        got, want = b.step(x, dt), ref.step(x, dt)
        assert torch.equal(got, want)
        pu = scale * torch.rand(rows, b.L, device=cuda_device, generator=g, dtype=b.dtype)
        assert torch.equal(b.step(pu, dt), ref.step(pu, dt))
    out, fat = b.step_env(x, dt)
    out2, fat2 = ref.step_env(x, dt)
    assert torch.equal(out, out2) and torch.equal(fat, fat2)
    if b.S == 1:
        xs = scale * torch.rand(37, rows, device=cuda_device, generator=g, dtype=b.dtype)
        assert torch.equal(b.step_many(xs, dt), ref.step_many(xs, dt))
    torch.cuda.synchronize()
    _bands_intact(arenas)
    assert torch.equal(b.cpf, ref.cpf) and torch.equal(b.dur, ref.dur)
    assert not torch.isnan(b.cpf).any() and not torch.isnan(b.dur).any()


# --------------------------------------------------------------------------------------------------
# K2x: one block per muscle -- wide muscles and rows of several muscles in ONE fused launch
# --------------------------------------------------------------------------------------------------
def test_wide_muscle_10000_steps_fused(cuda_device):
    """StandardMuscle(527.1 N) = 2,000 units over 10,000 steps in fused windows of 1,000 (muscle.py:144-147,
    :189-243 make such muscles first-class), FP32 mode, against the oracle: the north star's rtol 1e-4."""
    M, K, W, dt = 2, 10000, 1000, 0.02
    spec = MuscleSpec.standard(527.1)
    assert spec.n == 2000
    ora = OracleMuscles.standard(527.1, m=M)
    rng = np.random.default_rng(17)
    level = np.array([0.6, 0.95])
    b = MuscleBatch(spec, M, cuda_device, "f32")
    launches0 = b.launches
    ptf = spec.tables()["ptf"]
    for w0 in range(0, K, W):
        # piecewise-constant drive with rests, so that units fatigue, recover and cross zero capacity
        seg = np.where((np.arange(w0, w0 + W) // 700) % 3 == 2, 0.0, 1.0)[:, None] * level[None, :]
        seg = (seg * (0.9 + 0.1 * rng.random((W, M)))).astype(np.float32)
        got = b.step_many(torch.from_numpy(seg).to(cuda_device), dt)
        want = oracle_window(ora, seg, dt)
        close(got.cpu().numpy(), want, "f32", total_scale(spec), f"totals of window {w0 // W}")
        close(b.cpf.cpu().numpy(), ora.cpf, "f32", ptf, f"capacities after window {w0 // W}")
        close(b.current_forces.cpu().numpy(), ora.current_forces, "f32", ptf, f"forces after window {w0 // W}")
    assert b.launches - launches0 == K // W, "a window of a 2,000-unit muscle must be one launch"


@pytest.mark.parametrize("forces_n", [(9.0, 33.0, 21.0), (20.0, 32.0, 45.0, 32.0), (68.0, 20.0), (68.0, 824.0), (68.0, 300.0)])
def test_ragged_rows_fused_window(cuda_device, forces_n):
    """Rows of several muscles in ONE fused launch, with held controls and decimated totals, against per-muscle oracles.
    Short muscles run the warp-per-muscle kernel -- (20, 32, 45, 32 N) = a hopper leg's 74 / 120 / 169 / 120 units = 3 pairs per
    lane, (68 N, 20 N) = 256 and 74 units = 4 pairs -- rows with a longer muscle the block-per-muscle kernel: (68 N, 300 N) =
    256 and 1,137 units in 256-thread blocks, (68 N, 824 N) = 256 and 3,127 units in the 512-thread geometry."""
    specs = [MuscleSpec.standard(f) for f in forces_n]
    S, rows, K, dt, hold, tev = len(specs), 5, 96, 0.02, 4, 3
    b = MuscleBatch(specs, rows, cuda_device, "f32")
    oras = [OracleMuscles.standard(f, m=rows) for f in forces_n]
    rng = np.random.default_rng(23)
    ctrl = rng.random((K // hold, rows, S)).astype(np.float32)
    launches0 = b.launches
    got, fdec, masks = b.step_many(torch.from_numpy(ctrl.reshape(K // hold, rows * S)).to(cuda_device), dt, steps=K,
                                   hold=hold, totals_every=tev, forces_every=32, return_masks=True)
    assert b.launches - launches0 == 1
    want = np.empty((K, rows, S))
    fwant, mwant = [], []
    for k in range(K):
        recs = [o.step(ctrl[k // hold][:, i].astype(np.float64), dt) for i, o in enumerate(oras)]
        want[k] = np.stack([r.total for r in recs], axis=1)
        if (k + 1) % 32 == 0:
            fwant.append(np.concatenate([r.forces for r in recs], axis=1))
            mwant.append(np.concatenate([r.mask for r in recs], axis=1))
    got = got.cpu().numpy().reshape(K // tev, rows, S)
    for i, sp in enumerate(specs):
        close(got[:, :, i], want[tev - 1::tev, :, i], "f32", total_scale(sp), f"totals of muscle {i}")
    ptf = np.concatenate([sp.tables()["ptf"] for sp in specs])
    close(fdec.cpu().numpy(), np.stack(fwant), "f32", ptf, "decimated forces")
    assert np.array_equal(masks.cpu().numpy().astype(bool), np.stack(mwant)), "recruitment masks must be exact"
    close(b.cpf.cpu().numpy(), np.concatenate([o.cpf for o in oras], axis=1), "f32", ptf, "capacities")
    close(b.current_forces.cpu().numpy(), np.concatenate([o.current_forces for o in oras], axis=1), "f32", ptf, "last forces")


def test_wide_kernel_with_central_fatigue_and_potvin_fibers(cuda_device):
    """Potvin muscles wider than the block kernel's 480 units (central fatigue on: the adaptation recurrence)."""
    n, M, K, dt = 700, 3, 200, 0.02
    b = MuscleBatch(MuscleSpec.potvin(n), M, cuda_device, "f32")
    ora = OracleMuscles(n, M, "potvin")
    rng = np.random.default_rng(29)
    exc = (67.0 * rng.random((K, M))).astype(np.float32)
    got = b.step_many(torch.from_numpy(exc).to(cuda_device), dt)
    close(got.cpu().numpy(), oracle_window(ora, exc, dt), "f32", float(ora.t.ptf.sum()), "totals")
    close(b.cpf.cpu().numpy(), ora.cpf, "f32", ora.t.ptf, "capacities")
    np.testing.assert_allclose(b.dur.cpu().numpy(), ora.dur, rtol=1e-9, atol=1e-12)


def test_very_short_windows_use_the_persistent_prefetching_kernel(cuda_device):
    """Windows of <= 24 steps on a batch larger than the resident warps run k2w_f32's persistent variant (cp.async prefetch of
    the next muscle pair, direct state stores, red.add on the durations): against the oracle, over several windows, with an
    odd muscle count and units that rest, fatigue and recover across window boundaries."""
    n, M, K, dt = 40, 4801, 20, 0.02
    sm = torch.cuda.get_device_properties(cuda_device).multi_processor_count
    for kind in ("potvin", "standard"):
        spec = MuscleSpec.potvin(n) if kind == "potvin" else MuscleSpec(n, fibers=MuscleSpec.standard().fibers,
                                                                         pool=MuscleSpec.standard().pool,
                                                                         input_scale=67.0, output_divisor=1.0)
        b = MuscleBatch(spec, M, cuda_device, "f32")
        if (M + 15) // 16 > 2 * sm:
            assert b.fused_geometry(K)["blocks"] == 2 * sm, "expected the persistent geometry (resident blocks only)"
        ora = OracleMuscles(n, M, kind)
        if kind == "standard":
            ora.max_arb_output = 1.0                                  # (a PyMuscle-fibers muscle with unit output divisor)
        rng = np.random.default_rng(31)
        level = rng.random(M)
        for w in range(6):
            gate = 0.0 if w in (2, 3) else 1.0                        # two windows at rest: recovery / untouched units
            exc = (gate * level[None, :] * (0.8 + 0.2 * rng.random((K, M)))).astype(np.float32)
            if kind == "potvin":
                exc = exc * 67.0
            got = b.step_many(torch.from_numpy(exc).to(cuda_device), dt)
            want = oracle_window(ora, exc, dt)
            close(got.cpu().numpy(), want, "f32", float(ora.t.ptf.sum()), f"{kind} totals, window {w}")
            close(b.cpf.cpu().numpy(), ora.cpf, "f32", ora.t.ptf, f"{kind} capacities, window {w}")
            np.testing.assert_allclose(b.dur.cpu().numpy(), ora.dur, rtol=1e-9, atol=1e-12)
        if kind == "potvin":
            untouched = ora.cpf == ora.t.ptf                          # never recruited: must keep the exact float64 peak
            assert untouched.any() and np.array_equal(b.cpf.cpu().numpy()[untouched], ora.cpf[untouched])


@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_hopper_rows_3000_steps_single_and_fused(cuda_device, precision):
    """A hopper leg's four muscles (74 / 120 / 169 / 120 units; examples/envs/muscled_hopper.py:10-23) over 3,000 steps with
    bursts and rests: env steps through the flat-tiled streaming kernel (both modes) and, in FP32 mode, the same protocol as
    fused windows through the warp-per-muscle kernel -- outputs, fatigue readout, capacities against per-muscle oracles at
    the mode's tolerance."""
    forces_n, rows, K, dt = (20.0, 32.0, 45.0, 32.0), 3, 3000, 0.02
    specs = [MuscleSpec.standard(f) for f in forces_n]
    S = len(specs)
    rng = np.random.default_rng(41)
    base = rng.random((rows, S))
    gate = ((np.arange(K) // 400) % 3 != 2).astype(np.float64)          # 800 steps on, 400 at rest
    act = (gate[:, None, None] * base[None] * (0.85 + 0.15 * rng.random((K, rows, S)))).astype(np.float32)
    oras = [OracleMuscles.standard(f, m=rows) for f in forces_n]
    want = np.empty((K, rows, S))
    want_fat = np.empty((K, rows, S))
    for k in range(K):
        for i, o in enumerate(oras):
            want[k, :, i] = o.step(act[k, :, i].astype(np.float64), dt).total
            want_fat[k, :, i] = o.peripheral_fatigue()
    ptf = np.concatenate([sp.tables()["ptf"] for sp in specs])
    cpf_want = np.concatenate([o.cpf for o in oras], axis=1)

    b = MuscleBatch(specs, rows, cuda_device, precision)
    x = torch.from_numpy(act).to(cuda_device, b.dtype)
    outs, fats = [], []
    for k in range(K):
        o, f = b.step_env(x[k], dt)
        outs.append(o)
        fats.append(f)
    got = torch.stack(outs).double().cpu().numpy()
    for i, sp in enumerate(specs):
        close(got[:, :, i], want[:, :, i], precision, total_scale(sp), f"single-step totals, muscle {i}")
    close(torch.stack(fats).cpu().numpy(), want_fat, precision, 1.0, "fatigue readout")
    close(b.cpf.cpu().numpy(), cpf_want, precision, ptf, "capacities after 3,000 single steps")

    if precision == "f32":
        fb = MuscleBatch(specs, rows, cuda_device, "f32")
        l0 = fb.launches
        tot = torch.cat([fb.step_many(x[w:w + 500].reshape(500, rows * S), dt) for w in range(0, K, 500)])
        assert fb.launches - l0 == K // 500, "a window of ragged rows must be one launch"
        tot = tot.cpu().numpy().reshape(K, rows, S)
        for i, sp in enumerate(specs):
            close(tot[:, :, i], want[:, :, i], "f32", total_scale(sp), f"fused totals, muscle {i}")
        close(fb.cpf.cpu().numpy(), cpf_want, "f32", ptf, "capacities after 3,000 fused steps")
