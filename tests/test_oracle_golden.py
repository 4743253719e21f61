"""
The oracle (oracle/pymuscle_oracle.py) against (i) the known-answer values the
reference's own test-suite pins and (ii) golden trajectories generated from the
unmodified reference (tests/golden/make_golden.py).  CPU only.
"""
import numpy as np
import pytest

from conftest import golden, trajectory_names, traj_kwargs
from oracle.pymuscle_oracle import (OracleMuscle, OracleMuscles, PoolHyper, FibersHyper, UnitTables,
                                    force_to_motor_unit_count, recruitment_mask)

# The generating container and the GPU box run the same image, so the oracle
# normally reproduces the fixtures bit for bit; 1e-12 leaves room for a NumPy
# build that dispatches exp/pow to a different SIMD kernel.
RTOL = 1e-12


# ---- (i) values pinned by the reference's tests ------------------

def test_tables_match_reference():
    g = golden("tables")
    for n in (2, 33, 120, 340, 2000):
        t = UnitTables.build(n, PoolHyper(), FibersHyper(recovery=True))
        for key, arr in (("thr", t.thr), ("peak", t.peak), ("ptf", t.ptf), ("ct", t.ct), ("fat", t.fat), ("rec", t.rec)):
            np.testing.assert_allclose(arr, g[f"{key}_{n}"], rtol=RTOL, atol=0)
    assert t.max_excitation == 67.0                     # tests/test_..._motor_neuron_pool.py:16
    for mf, cnt in g["counts_for_force"]:
        assert force_to_motor_unit_count(float(mf), 0.0123) == int(cnt)
    assert force_to_motor_unit_count(32.0, 0.0123) == 120       # tests/test_standard_muscle.py:11
    assert force_to_motor_unit_count(90.0, 0.0123) == 340       # :21
    assert OracleMuscles.standard(32.0).max_arb_output == pytest.approx(2609.0309)   # :14-15
    assert OracleMuscles.standard(90.0).max_arb_output == pytest.approx(7338.29062)  # :22-23


def test_pool_known_answers():
    # tests/test_potvin_fuglevand_2017_motor_neuron_pool.py:36-81
    for e, want in ((0.0, 0.0), (40.0, 3503.58881), (67.0, 3915.06787), (107.0, 3915.06787)):
        o = OracleMuscles(120)
        _, rates = o.pool_step(np.full((1, 120), e), 1.0)
        assert np.sum(rates) == pytest.approx(want)
    o = OracleMuscles(120)
    for _ in range(11):
        _, rates = o.pool_step(np.full((1, 120), 67.0), 1.0)
    assert np.sum(rates) == pytest.approx(3749.91061)
    o = OracleMuscles(120, apply_central_fatigue=False)
    first = np.sum(o.pool_step(np.full((1, 120), 67.0), 1.0)[1])
    for _ in range(10):
        nxt = np.sum(o.pool_step(np.full((1, 120), 67.0), 1.0)[1])
    assert nxt == first


@pytest.mark.parametrize("kind", ["potvin", "standard"])
def test_fibers_known_answers(kind):
    # tests/test_potvin_fuglevand_2017_muscle_fibers.py:32-56, tests/test_pymuscle_fibers.py:35-60
    for rate, want in ((0.0, 0.0), (40.0, 2586.7530897), (67.0, 2609.0308816), (107.0, 2609.0308816)):
        o = OracleMuscles(120, kind=kind)
        _, total = o.fibers_step(np.full((1, 120), rate), 1.0)
        assert total[0] == pytest.approx(want)


def test_pymuscle_fibers_fatigue_invariants():
    # tests/test_pymuscle_fibers.py:63-92
    o = OracleMuscles(120, kind="standard")
    before = o.cpf.copy()
    for _ in range(100):
        o.fibers_step(np.zeros((1, 120)), 1.0)
    assert np.array_equal(before, o.cpf)
    for _ in range(100):
        o.fibers_step(np.full((1, 120), 67.0), 1.0)
    assert np.all(before > o.cpf)
    o = OracleMuscles(120, kind="standard", apply_peripheral_fatigue=False)
    for _ in range(100):
        o.fibers_step(np.full((1, 120), 67.0), 1.0)
    assert np.array_equal(before, o.cpf)


def test_potvin_muscle_known_answers():
    # tests/test_potvin_fuglevand_muscle.py:26-52
    for e, want in ((0.0, 0.0), (40.0, 1311.86896), (67.0, 2215.98114), (107.0, 2215.98114)):
        assert OracleMuscles(120).step(e, 1.0).total[0] == pytest.approx(want)
        assert OracleMuscle(120).step(e, 1.0) == pytest.approx(want)


def test_standard_muscle_known_answers():
    # tests/test_standard_muscle.py:36-114
    for x, want in ((0.0, 0.0), (0.5, 0.39096348), (1.0, 0.84935028), (41.0, 0.84935028)):
        assert OracleMuscles.standard(32.0).step(x, 1.0).total[0] == pytest.approx(want)
        assert OracleMuscle(120, "standard").step(x, 1.0) == pytest.approx(want)
    o = OracleMuscles.standard(32.0, apply_peripheral_fatigue=False)
    for _ in range(100):
        out = o.step(0.5, 1.0).total[0]
    assert out == pytest.approx(0.39096348) and o.peripheral_fatigue()[0] == 0.0
    o = OracleMuscles.standard(32.0)
    for _ in range(100):
        out = o.step(0.0, 1.0).total[0]
    assert out == pytest.approx(0.0) and o.peripheral_fatigue()[0] == 0.0
    for _ in range(100):
        out = o.step(0.5, 1.0).total[0]
    assert out == pytest.approx(0.29151827)
    assert o.peripheral_fatigue()[0] == pytest.approx(0.18028181)


# ---- (ii) golden trajectories from the live reference -----------------------------------

@pytest.mark.parametrize("name", trajectory_names())
def test_trajectory(name):
    g = golden("traj_" + name)
    exc, dt, n = g["exc"], float(g["dt"]), int(g["n"])
    K, m = exc.shape[:2]
    o = OracleMuscles(n, m, str(g["kind"]), **traj_kwargs(g))
    cps = list(g["checkpoints"])
    totals = np.zeros((K, m))
    for k in range(K):
        rec = o.step(np.asarray(exc[k], dtype=np.float64), dt)
        totals[k] = rec.total
        if k in cps:
            i = cps.index(k)
            np.testing.assert_allclose(rec.forces, g["forces"][i], rtol=RTOL, atol=1e-300)
            np.testing.assert_allclose(o.cpf, g["cpf"][i], rtol=RTOL, atol=1e-300)
            np.testing.assert_allclose(o.dur, g["dur"][i], rtol=RTOL, atol=0)
    np.testing.assert_allclose(totals, g["totals"], rtol=RTOL, atol=1e-300)


def test_single_muscle_form_matches_batched():
    """The reference-shaped OracleMuscle (what the CPU baseline times) and the
    batched OracleMuscles are the same arithmetic."""
    g = golden("traj_standard_random")
    exc = g["exc"][:300]
    one = OracleMuscle(120, "standard")
    want = g["totals"][:300, 0]
    got = np.array([one.step(float(e), 0.02) for e in exc[:, 0]])
    np.testing.assert_allclose(got, want, rtol=RTOL)


def test_mask_probes():
    g = golden("mask_probes")
    t, h = UnitTables.build(120), PoolHyper()
    for probes, packed, scale in ((g["probes64"], g["mask64"], 1.0), (g["probes32"], g["mask32"], 1.0),
                                  (g["xprobes32"], g["xmask32"], 67.0)):
        e = probes.astype(np.float64)[:, None] * scale
        got = recruitment_mask(np.broadcast_to(e, (len(probes), 120)), t, h)
        assert np.array_equal(np.packbits(got, axis=1), packed)


def test_halves():
    g = golden("halves")
    o = OracleMuscles(120)
    rates = np.stack([o.pool_step(e.astype(np.float64)[None], 0.5)[1][0] for e in g["pool_exc"]])
    np.testing.assert_allclose(rates, g["pool_rates"], rtol=RTOL, atol=1e-300)
    np.testing.assert_allclose(o.dur[0], g["pool_dur"], rtol=RTOL)
    for name, kind in (("fib", "potvin"), ("pmf", "standard")):
        o = OracleMuscles(120, kind=kind)
        tot = [o.fibers_step(r.astype(np.float64)[None], 0.5)[1][0] for r in g[f"{name}_rates"]]
        np.testing.assert_allclose(tot, g[f"{name}_totals"], rtol=RTOL)
        np.testing.assert_allclose(o.cpf[0], g[f"{name}_cpf"], rtol=RTOL, atol=1e-300)
        np.testing.assert_allclose(o.current_forces[0], g[f"{name}_forces_last"], rtol=RTOL, atol=1e-300)


def test_arm():
    g = golden("arm")
    counts = [force_to_motor_unit_count(float(f), 0.0123) for f in g["max_forces"]]
    assert counts == list(g["counts"]) and sum(counts) == 35032
    muscles = [OracleMuscles(c, 1, "standard") for c in counts]
    for k in range(g["exc"].shape[0]):
        for j, mu in enumerate(muscles):
            got = mu.step(float(g["exc"][k, j]), float(g["dt"])).total[0]
            assert got == pytest.approx(g["totals"][k, j], rel=RTOL, abs=1e-300)
    np.testing.assert_allclose(np.concatenate([mu.cpf[0] for mu in muscles]), g["cpf"], rtol=RTOL)


def test_readme_exact_values():
    """SURVEY.md Appendix D (literal 0.6 / 40.0 inputs), via the reference-shaped oracle."""
    g = golden("readme_exact")
    assert g["standard_totals"][0] == pytest.approx(0.5059615150545111, rel=1e-13)
    assert g["standard_totals"][9999] == pytest.approx(0.1435819929896607, rel=1e-13)
    assert float(g["standard_fatigue"]) == pytest.approx(0.5413389785180036, rel=1e-13)
    assert g["potvin_totals"][0] == pytest.approx(1311.8689626905448, rel=1e-13)
    assert g["potvin_totals"][9999] == pytest.approx(403.39412010089995, rel=1e-13)
    assert float(g["rest_then_half_output"]) == pytest.approx(0.291518275974285, rel=1e-13)
    assert float(g["rest_then_half_fatigue"]) == pytest.approx(0.1802818182250192, rel=1e-13)
    m = OracleMuscle(120, "standard")
    got = np.array([m.step(0.6, 1 / 50.0) for _ in range(2000)])
    np.testing.assert_allclose(got, g["standard_totals"][:2000], rtol=RTOL)
