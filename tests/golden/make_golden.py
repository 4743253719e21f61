#!/usr/bin/env python
"""
Generate golden input/output vectors from the UNMODIFIED reference
(iandanforth/pymuscle v0.1.2 imported from /root/reference) and, with --check,
prove the NumPy oracle (oracle/pymuscle_oracle.py) bit-identical to it.

    python tests/golden/make_golden.py            # (re)write tests/golden/*.npz
    python tests/golden/make_golden.py --check    # oracle == live reference, bit for bit

Runs only in the build container: the reference does not travel to the GPU
box, the .npz files do.  Every scenario is driven through the reference's
public API (Muscle.step / Pool.step / Fibers.step / current_forces) and the
private state arrays are read (never written) for the capacity / duration
checkpoints.
"""
import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("PYMUSCLE_REFERENCE", "/root/reference")

CHECKPOINTS_10K = [0, 1, 99, 999, 2499, 4999, 7499, 9999]
ARM_FORCES = [round(68.0 * (824.0 / 68.0) ** (i / 29.0), 1) for i in range(30)]


def _ref():
    sys.path.insert(0, REF)
    import pymuscle  # noqa
    return pymuscle


def protocol_ramp_hold_rest(steps, e_max):
    """Config #4 protocol: ramp 0->e_max over the first 20 %, hold 40 %, rest 40 %."""
    k = np.arange(steps)
    a, b = steps // 5, (steps * 3) // 5
    e = np.where(k < a, k / float(a), np.where(k < b, 1.0, 0.0))
    return (e[:, None] * np.asarray(e_max)[None, :])


def f32(x):
    """Round to float32-representable values (so FP32-mode kernels see the same inputs)."""
    return np.asarray(x, dtype=np.float32)


def run_reference(make_muscle, exc, dt, checkpoints, per_unit=False, fibers_attr="_fibers", pool_attr="_pool"):
    """exc: [K, m] (per muscle) or [K, m, n] (per unit).  Returns dict of arrays."""
    K, m = exc.shape[0], exc.shape[1]
    muscles = [make_muscle() for _ in range(m)]
    n = muscles[0].motor_unit_count
    totals = np.zeros((K, m))
    cps = sorted(set(c for c in checkpoints if c < K))
    forces = np.zeros((len(cps), m, n))
    cpf = np.zeros((len(cps), m, n))
    dur = np.zeros((len(cps), m, n))
    for k in range(K):
        for j, mu in enumerate(muscles):
            if per_unit:
                e = np.array(exc[k, j], dtype=np.float64)      # fresh array: StandardMuscle scales in place
            else:
                e = float(exc[k, j])
            totals[k, j] = mu.step(e, dt)
        if k in cps:
            i = cps.index(k)
            for j, mu in enumerate(muscles):
                forces[i, j] = mu.current_forces
                cpf[i, j] = getattr(mu, fibers_attr)._current_peak_forces
                dur[i, j] = getattr(mu, pool_attr)._recruitment_durations
    return dict(totals=totals, checkpoints=np.array(cps), forces=forces, cpf=cpf, dur=dur)


def run_oracle(kind, n, exc, dt, checkpoints, per_unit=False, **kw):
    sys.path.insert(0, ROOT)
    from oracle.pymuscle_oracle import OracleMuscles
    K, m = exc.shape[0], exc.shape[1]
    o = OracleMuscles(n, m, kind, **kw)
    totals = np.zeros((K, m))
    cps = sorted(set(c for c in checkpoints if c < K))
    forces = np.zeros((len(cps), m, n))
    cpf = np.zeros((len(cps), m, n))
    dur = np.zeros((len(cps), m, n))
    for k in range(K):
        rec = o.step(np.asarray(exc[k], dtype=np.float64), dt)
        totals[k] = rec.total
        if k in cps:
            i = cps.index(k)
            forces[i], cpf[i], dur[i] = rec.forces, o.cpf, o.dur
    return dict(totals=totals, checkpoints=np.array(cps), forces=forces, cpf=cpf, dur=dur)


def scenarios(pm):
    """name -> (kind, n, kwargs, make_muscle, exc, dt, checkpoints, per_unit)"""
    rng = np.random.default_rng(2017)
    S = {}
    # --- BASELINE config #1: README loop ---------------------------------
    S["standard_const06"] = ("standard", 120, {}, lambda: pm.StandardMuscle(),
                             f32(np.full((10000, 1), 0.6)), 0.02, CHECKPOINTS_10K, False)
    # --- examples/minimal-physio-example.py shape ------------------------
    S["potvin_const40"] = ("potvin", 120, {}, lambda: pm.PotvinFuglevandMuscle(120),
                           f32(np.full((10000, 1), 40.0)), 0.02, CHECKPOINTS_10K, False)
    # --- config #4 protocol (ramp / hold / rest), 4 muscles --------------
    u = rng.random(4)
    S["potvin_ramp"] = ("potvin", 120, {}, lambda: pm.PotvinFuglevandMuscle(120),
                        f32(protocol_ramp_hold_rest(10000, 67.0 * (0.5 + 0.5 * u))), 0.02, CHECKPOINTS_10K, False)
    S["standard_ramp"] = ("standard", 120, {}, lambda: pm.StandardMuscle(),
                          f32(protocol_ramp_hold_rest(10000, 0.5 + 0.5 * u)), 0.02, CHECKPOINTS_10K, False)
    # --- per-step random excitation (configs #3/#5 style) ----------------
    S["potvin_random"] = ("potvin", 120, {}, lambda: pm.PotvinFuglevandMuscle(120),
                          f32(67.0 * rng.random((2000, 4))), 0.02, [0, 1, 499, 1999], False)
    S["standard_random"] = ("standard", 120, {}, lambda: pm.StandardMuscle(),
                            f32(rng.random((2000, 4))), 0.02, [0, 1, 499, 1999], False)
    # --- on/off bursts: exercises recovery and the zero clamp ------------
    burst = (rng.random((3000, 3)) < 0.5) * rng.random((3000, 3))
    burst = np.repeat(burst[::50], 50, axis=0)                # 1-second on/off blocks
    S["standard_bursts_dt1"] = ("standard", 120, {}, lambda: pm.StandardMuscle(),
                                f32(burst), 1.0, [0, 49, 999, 2999], False)
    # --- other unit counts ------------------------------------------------
    S["standard_n340"] = ("standard", 340, {}, lambda: pm.StandardMuscle(90.0),
                          f32(rng.random((500, 2))), 0.02, [0, 499], False)
    S["standard_n2000"] = ("standard", 2000, {}, lambda: pm.StandardMuscle(527.1),
                           f32(rng.random((200, 2))), 0.02, [0, 199], False)
    S["potvin_n33"] = ("potvin", 33, {}, lambda: pm.PotvinFuglevandMuscle(33),
                       f32(67.0 * rng.random((300, 3))), 0.05, [0, 299], False)
    S["potvin_n2"] = ("potvin", 2, {}, lambda: pm.PotvinFuglevandMuscle(2),
                      f32(67.0 * rng.random((100, 2))), 0.05, [0, 99], False)
    # --- fatigue switches -------------------------------------------------
    S["potvin_nocentral"] = ("potvin", 120, dict(apply_central_fatigue=False),
                             lambda: pm.PotvinFuglevandMuscle(120, apply_central_fatigue=False),
                             f32(67.0 * rng.random((500, 2))), 0.02, [0, 499], False)
    S["potvin_noperiph"] = ("potvin", 120, dict(apply_peripheral_fatigue=False),
                            lambda: pm.PotvinFuglevandMuscle(120, apply_peripheral_fatigue=False),
                            f32(67.0 * rng.random((500, 2))), 0.02, [0, 499], False)
    S["standard_central"] = ("standard", 120, dict(apply_central_fatigue=True),
                             lambda: pm.StandardMuscle(apply_central_fatigue=True),
                             f32(rng.random((500, 2))), 0.02, [0, 499], False)
    S["standard_noperiph"] = ("standard", 120, dict(apply_peripheral_fatigue=False),
                              lambda: pm.StandardMuscle(apply_peripheral_fatigue=False),
                              f32(rng.random((300, 2))), 0.02, [0, 299], False)
    # --- per-unit excitation vectors (Muscle.step accepts ndarray[n]) -----
    S["potvin_perunit"] = ("potvin", 120, {}, lambda: pm.PotvinFuglevandMuscle(120),
                           f32(67.0 * rng.random((200, 2, 120))), 0.02, [0, 199], True)
    S["standard_perunit"] = ("standard", 120, {}, lambda: pm.StandardMuscle(),
                             f32(rng.random((200, 2, 120))), 0.02, [0, 199], True)
    return S


def mask_probes(pm):
    """Excitations within a few ulps of every threshold (float64 and float32
    grids) and the reference's recruitment decision for each."""
    pool = pm.PotvinFuglevand2017MotorNeuronPool(120)
    thr = pool._recruitment_thresholds
    offs = [0, 1, 2, 3, 10, 100, 1000]
    probes64 = []
    for o in offs:
        up, dn = thr.copy(), thr.copy()
        for _ in range(o if o <= 10 else 0):
            up, dn = np.nextafter(up, np.inf), np.nextafter(dn, -np.inf)
        if o > 10:
            up = thr * (1 + o * 2.0 ** -52)
            dn = thr * (1 - o * 2.0 ** -52)
        probes64 += [up, dn]
    probes64 = np.concatenate(probes64)
    t32 = thr.astype(np.float32)
    probes32 = []
    for o in range(-4, 5):
        p = t32.copy()
        for _ in range(abs(o)):
            p = np.nextafter(p, np.float32(np.inf if o > 0 else -np.inf))
        probes32.append(p)
    probes32 = np.concatenate(probes32)

    def decide(e_scalar):
        r = pool._inner_calc_firing_rates(np.full(120, e_scalar), thr, pool._firing_gain,
                                          pool._min_firing_rate, pool._peak_firing_rates)
        return r > 0

    m64 = np.stack([decide(float(e)) for e in probes64])
    m32 = np.stack([decide(float(e)) for e in probes32])
    # StandardMuscle inputs are 0..1 and get multiplied by max_excitation first (muscle.py:275)
    x32 = []
    for o in range(-4, 5):
        p = (thr / 67.0).astype(np.float32)
        for _ in range(abs(o)):
            p = np.nextafter(p, np.float32(np.inf if o > 0 else -np.inf))
        x32.append(p)
    x32 = np.concatenate(x32)
    mx32 = np.stack([decide(float(np.float64(x) * pool.max_excitation)) for x in x32])
    return dict(probes64=probes64, mask64=np.packbits(m64, axis=1),
                probes32=probes32, mask32=np.packbits(m32, axis=1),
                xprobes32=x32, xmask32=np.packbits(mx32, axis=1))


def halves(pm):
    """Pool.step and Fibers.step on their own (the Model protocol, model.py:4-18)."""
    rng = np.random.default_rng(93)
    out = {}
    pool = pm.PotvinFuglevand2017MotorNeuronPool(120)
    exc = f32(67.0 * rng.random((50, 120)))
    out["pool_exc"] = exc
    out["pool_rates"] = np.stack([pool.step(np.array(e, dtype=np.float64), 0.5) for e in exc])
    out["pool_dur"] = pool._recruitment_durations.copy()
    for name, cls in (("fib", pm.PotvinFuglevand2017MuscleFibers), ("pmf", pm.PyMuscleFibers)):
        f = cls(120)
        rates = f32(35.0 * rng.random((50, 120)) * (rng.random((50, 120)) < 0.7))
        tot, forces = [], []
        for r in rates:
            tot.append(f.step(np.array(r, dtype=np.float64), 0.5))
            forces.append(f.current_forces.copy())
        out[f"{name}_rates"] = rates
        out[f"{name}_totals"] = np.array(tot)
        out[f"{name}_forces_last"] = forces[-1]
        out[f"{name}_cpf"] = f._current_peak_forces.copy()
    return out


def arm(pm):
    """Config #3: 30 StandardMuscles of increasing size, one agent, 20 random steps."""
    rng = np.random.default_rng(3)
    muscles = [pm.StandardMuscle(max_force=f) for f in ARM_FORCES]
    counts = np.array([m.motor_unit_count for m in muscles])
    exc = f32(rng.random((20, 30)))
    totals = np.zeros((20, 30))
    for k in range(20):
        for j, m in enumerate(muscles):
            totals[k, j] = m.step(float(exc[k, j]), 0.02)
    cpf = np.concatenate([m._fibers._current_peak_forces for m in muscles])
    forces = np.concatenate([m.current_forces for m in muscles])
    return dict(max_forces=np.array(ARM_FORCES), counts=counts, exc=exc, totals=totals,
                cpf=cpf, forces=forces, dt=0.02)


def readme_exact(pm):
    """The README loop with the literal Python float 0.6 (not float32-rounded) and the
    minimal-physio example with 40.0: the values SURVEY.md Appendix D quotes."""
    m = pm.StandardMuscle()
    tot = np.array([m.step(0.6, 1 / 50.0) for _ in range(10000)])
    out = dict(standard_totals=tot, standard_fatigue=np.array(m.get_peripheral_fatigue()),
               standard_cpf=m._fibers._current_peak_forces.copy(), standard_forces=m.current_forces.copy())
    m = pm.PotvinFuglevandMuscle(120)
    tot = np.array([m.step(40.0, 0.02) for _ in range(10000)])
    out.update(potvin_totals=tot, potvin_cpf=m._fibers._current_peak_forces.copy(),
               potvin_dur=m._pool._recruitment_durations.copy())
    m = pm.StandardMuscle(32.0)
    for _ in range(100):
        m.step(0.0, 1.0)
    for _ in range(100):
        o = m.step(0.5, 1.0)
    out.update(rest_then_half_output=np.array(o), rest_then_half_fatigue=np.array(m.get_peripheral_fatigue()))
    return out


def tables(pm):
    out = {}
    for n in (2, 33, 120, 340, 2000):
        p = pm.PotvinFuglevand2017MotorNeuronPool(n)
        f = pm.PyMuscleFibers(n)
        out[f"thr_{n}"] = p._recruitment_thresholds
        out[f"peak_{n}"] = p._peak_firing_rates
        out[f"ptf_{n}"] = f._peak_twitch_forces
        out[f"ct_{n}"] = f._contraction_times
        out[f"fat_{n}"] = f._nominal_fatigabilities
        out[f"rec_{n}"] = f._recovery_rates
    out["counts_for_force"] = np.array([[mf, pm.StandardMuscle.force_to_motor_unit_count(mf, 0.0123)]
                                        for mf in (5.0, 32.0, 90.0, 527.1, 824.0)])
    out["max_arb_output_120"] = np.array(pm.StandardMuscle().max_arb_output)
    return out


def hill(pm):
    """Hill-type helper curves (pymuscle/hill_type.py) on random lengths."""
    from pymuscle import hill_type as h
    rng = np.random.default_rng(1)
    rest = 0.5 + rng.random(64)
    cur = rest * (0.7 + 0.8 * rng.random(64))
    prev = cur * (0.95 + 0.1 * rng.random(64))
    return dict(rest=rest, cur=cur, prev=prev, dt=0.02,
                fl=h.contractile_element_force_length_curve(rest, cur),
                fv=h.contractile_element_force_velocity_curve(rest, cur, prev, 0.02))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true", help="compare oracle with the live reference bit for bit")
    args = ap.parse_args()
    pm = _ref()
    sc = scenarios(pm)
    if args.check:
        bad = 0
        for name, (kind, n, kw, make, exc, dt, cps, pu) in sc.items():
            r = run_reference(make, exc, dt, cps, pu)
            o = run_oracle(kind, n, exc, dt, cps, pu, **kw)
            for key in ("totals", "forces", "cpf", "dur"):
                same = np.array_equal(r[key], o[key])
                worst = np.max(np.abs(r[key] - o[key]) / np.maximum(np.abs(r[key]), 1e-300)) if not same else 0.0
                print(f"{name:22s} {key:7s} bit-identical={same} worst_rel={worst:.2e}")
                bad += (not same)
        print("ORACLE == REFERENCE (bitwise)" if bad == 0 else f"{bad} arrays differ")
        return 1 if bad else 0
    for name, (kind, n, kw, make, exc, dt, cps, pu) in sc.items():
        r = run_reference(make, exc, dt, cps, pu)
        np.savez_compressed(os.path.join(HERE, f"traj_{name}.npz"), kind=kind, n=n, dt=dt, exc=exc,
                            per_unit=pu, **{f"kw_{k}": v for k, v in kw.items()}, **r)
        print("wrote", name, r["totals"].shape)
    np.savez_compressed(os.path.join(HERE, "mask_probes.npz"), **mask_probes(pm))
    np.savez_compressed(os.path.join(HERE, "halves.npz"), **halves(pm))
    np.savez_compressed(os.path.join(HERE, "arm.npz"), **arm(pm))
    np.savez_compressed(os.path.join(HERE, "tables.npz"), **tables(pm))
    np.savez_compressed(os.path.join(HERE, "readme_exact.npz"), **readme_exact(pm))
    np.savez_compressed(os.path.join(HERE, "hill.npz"), **hill(pm))
    print("numpy", np.__version__)
    return 0


if __name__ == "__main__":
    sys.exit(main())
