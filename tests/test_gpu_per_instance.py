"""
Per-instance hyper-parameters (domain randomisation): every row of a batch has
its own packed parameter table.  Checked against the oracle with the same per-row hyper-parameters
-- the reference equivalent is a list of differently constructed Muscle objects.
"""
from dataclasses import replace

import numpy as np
import pytest
import torch

from oracle.pymuscle_oracle import FibersHyper, OracleMuscles, PoolHyper
from pymuscle_b200 import MuscleBatch, MuscleSpec
from pymuscle_b200.tables import pack_rows_per_instance
from pymuscle_b200 import _native as nat
from test_gpu_parity import close

pytestmark = pytest.mark.gpu


def randomised_rows(n, rows, kind, seed):
    """(specs, oracles): per-row recruitment range, peak rates, adaptation, twitch range, contraction
    times and fatigue rates; launch scalars untouched."""
    rng = np.random.default_rng(seed)
    specs, oras = [], []
    for _ in range(rows):
        pool_kw = dict(max_recruitment_threshold=float(rng.uniform(30, 70)),
                       max_firing_rate_first_unit=float(rng.uniform(30, 40)),
                       max_firing_rate_last_unit=float(rng.uniform(20, 28)),
                       derecruitment_delta=float(rng.uniform(1, 3)),
                       adaptation_magnitude=float(rng.uniform(0.4, 0.8)))
        fib_kw = dict(max_contraction_time=float(rng.uniform(70, 110)),
                      contraction_time_range=float(rng.uniform(2.5, 3.5)),
                      max_fatigue_rate=float(rng.uniform(0.015, 0.03)),
                      fatigability_range=float(rng.uniform(120, 240)),
                      contraction_time_change_ratio=float(rng.uniform(0.2, 0.5)))
        if kind == "potvin":
            fib_kw["max_twitch_amplitude"] = float(rng.uniform(60, 140))     # (would change a StandardMuscle's divisor)
            base = MuscleSpec.potvin(n)
        else:
            # a StandardMuscle's 0..1 input scaling (max_excitation, muscle.py:275) and output divisor are
            # per-launch scalars: keep what they depend on fixed, randomise the rest
            del pool_kw["max_recruitment_threshold"], pool_kw["max_firing_rate_last_unit"]
            base = MuscleSpec.standard()
            assert base.n == n
        specs.append(replace(base, pool=replace(base.pool, **pool_kw), fibers=replace(base.fibers, **fib_kw)))
        central = kind == "potvin"
        oras.append(OracleMuscles(n, 1, kind, pool=PoolHyper(apply_fatigue=central, **pool_kw),
                                  fib=FibersHyper(recovery=(kind == "standard"), **fib_kw)))
    return specs, oras


@pytest.mark.parametrize("precision", ["f32", "f64"])
@pytest.mark.parametrize("kind,n", [("potvin", 120), ("standard", 120), ("potvin", 33), ("potvin", 340)])
def test_single_and_fused_steps_with_per_row_tables(cuda_device, kind, n, precision):
    if kind == "standard" and n != 120:
        pytest.skip("StandardMuscle() has 120 units")
    rows, K, dt = 11, 70, 0.02
    specs, oras = randomised_rows(n, rows, kind, seed=n)
    scale = 1.0 if kind == "standard" else 80.0            # beyond max_excitation for some rows on purpose
    rng = np.random.default_rng(1)
    exc = (scale * rng.random((K + 3, rows))).astype(np.float32)
    b = MuscleBatch(specs[0], rows, cuda_device, precision, row_specs=specs)
    assert b.cfg.param_rows == rows
    x = torch.from_numpy(exc).to(cuda_device, b.dtype)
    got = [b.step(x[k], dt).cpu().numpy() for k in range(3)]                 # K1 (generic kernel, per-row tables)
    got_forces = b.current_forces.cpu().numpy()
    win = b.step_many(x[3:], dt).cpu().numpy()                                # K2 (k2u_f32 / block kernel, U = 1)
    totals = np.concatenate([np.stack(got), win])
    want = np.zeros_like(totals, dtype=np.float64)
    for r, o in enumerate(oras):
        for k in range(K + 3):
            rec = o.step(exc[k, r:r + 1].astype(np.float64), dt)
            want[k, r] = rec.total[0]
            if k == 2:
                close(got_forces[r], rec.forces[0], precision, o.t.ptf, f"forces after 3 single steps, row {r}")
        close(b.cpf[r].cpu().numpy(), o.cpf[0], precision, o.t.ptf, f"capacities, row {r}")
        np.testing.assert_allclose(b.dur[r].cpu().numpy(), o.dur[0], rtol=1e-9, atol=1e-12)
    out_scale = max(float(np.sum(o.t.ptf)) for o in oras) if kind == "potvin" else 1.0
    close(totals, want, precision, out_scale, "totals")


def test_rows_really_differ_and_launch_scalars_are_checked(cuda_device):
    specs, _ = randomised_rows(120, 4, "potvin", seed=9)
    b = MuscleBatch(specs[0], 4, cuda_device, "f32", row_specs=specs)
    t = b.step(torch.full((4,), 40.0, device=cuda_device), 0.02).cpu().numpy()
    assert len(set(np.round(t, 3))) == 4                                       # four different muscles
    bad = list(specs)
    bad[2] = replace(bad[2], pool=replace(bad[2].pool, firing_gain=1.5))
    with pytest.raises(ValueError, match="per-launch scalar"):
        pack_rows_per_instance(bad, nat.MB_F32)
    with pytest.raises(ValueError):
        MuscleBatch(specs[0], 5, cuda_device, "f32", row_specs=specs)
