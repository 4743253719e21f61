"""
CPU oracle for the PyMuscle per-step hot path.  TEST INFRASTRUCTURE ONLY.

This module is a NumPy (float64) restatement of the algorithm in
iandanforth/pymuscle v0.1.2 for the path

    Muscle.step -> PotvinFuglevand2017MotorNeuronPool.step
                -> PotvinFuglevand2017MuscleFibers.step / PyMuscleFibers.step

It exists so that the CUDA kernels in ``pymuscle_b200`` can be checked against
the reference's arithmetic on the GPU box, where ``/root/reference`` does not
exist.  It is NOT part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import it.  ``pymuscle_b200`` never does.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks it against
(i) every known-answer value in the reference's own test-suite and (ii) golden
trajectories produced by importing the unmodified reference in the build
container (``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``); in the
build container ``tests/golden/make_golden.py --check`` additionally proves the
oracle bit-identical to the live reference.

All citations are ``file:line`` into the reference tree, with the shorthand

    pool:    pymuscle/potvin_fuglevand_2017_motor_neuron_pool.py
    fibers:  pymuscle/potvin_fuglevand_2017_muscle_fibers.py
    pmf:     pymuscle/pymuscle_fibers.py
    muscle:  pymuscle/muscle.py

Unlike the reference (one Python object per muscle, arrays of shape [n]) every
function here works on arrays whose LAST axis is the motor-unit axis, so a
batch of M identical muscles is a [M, n] array.  NumPy ufuncs are elementwise,
so each row goes through exactly the reference's sequence of IEEE operations.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np

# --------------------------------------------------------------------------
# Hyper-parameters (defaults = the reference's constructor defaults)
# --------------------------------------------------------------------------


@dataclass(frozen=True)
class PoolHyper:
    """Constructor arguments of the motor-neuron pool, pool:53-66."""
    max_recruitment_threshold: float = 50
    firing_gain: float = 1.0
    min_firing_rate: float = 8
    max_firing_rate_first_unit: float = 35
    max_firing_rate_last_unit: float = 25
    derecruitment_delta: float = 2
    adaptation_magnitude: float = 0.67
    adaptation_time_constant: float = 22.0
    max_duration: float = 20000.0
    apply_fatigue: bool = True


@dataclass(frozen=True)
class FibersHyper:
    """Constructor arguments of the fibers models, fibers:46-56, pmf:55-60."""
    max_twitch_amplitude: float = 100
    max_contraction_time: float = 90
    contraction_time_range: float = 3
    max_fatigue_rate: float = 0.0225
    fatigability_range: float = 180
    contraction_time_change_ratio: float = 0.379
    apply_fatigue: bool = True
    recovery: bool = False          # True = PyMuscleFibers (pmf), False = Potvin fibers


# --------------------------------------------------------------------------
# Per-unit tables (built once)
# --------------------------------------------------------------------------

def _unit_indices(n: int) -> np.ndarray:
    return np.arange(1, n + 1)


def recruitment_thresholds(n: int, rr) -> np.ndarray:
    """thr_i = exp(ln(RR) * (i-1) / (n-1)) -- pool:263-281."""
    idx = _unit_indices(n)
    return np.exp((np.log(rr) * (idx - 1)) / (n - 1))


def peak_firing_rates(first, last, rr, thr: np.ndarray) -> np.ndarray:
    """Linear in the threshold between R1 and Rn -- pool:235-261."""
    span = first - last
    return first - (span * ((thr - thr[0]) / (rr - thr[0])))


def max_excitation(h: PoolHyper):
    """Excitation at which the last unit reaches its peak rate -- pool:96-99."""
    return h.max_recruitment_threshold + \
        (h.max_firing_rate_last_unit - h.min_firing_rate) / h.firing_gain


def peak_twitch_forces(n: int, rp) -> np.ndarray:
    """ptf_i = exp(ln(RP) * (i-1) / (n-1)) -- fibers:163-179."""
    idx = _unit_indices(n)
    return np.exp((np.log(rp) * (idx - 1)) / (n - 1))


def contraction_times(rp, t_l, rt, ptf: np.ndarray) -> np.ndarray:
    """ct_i = tL * (1/ptf_i)^(1/scale), scale = ln RP / ln rt -- fibers:127-161."""
    scale = np.log(rp) / np.log(rt)
    return t_l * np.power(1 / ptf, 1 / scale)


def nominal_fatigabilities(n: int, fat_range, max_rate, ptf: np.ndarray) -> np.ndarray:
    """fat_i = exp((ln FR/(n-1)) (i-1)) * (maxFat/FR) * ptf_i -- fibers:181-209."""
    idx = _unit_indices(n)
    curve = np.exp((np.log(fat_range) / (n - 1)) * (idx - 1))
    return curve * (max_rate / fat_range) * ptf


def recovery_rates(n: int, max_fatigue_rate, fat: np.ndarray, ptf: np.ndarray) -> np.ndarray:
    """PyMuscleFibers recovery table: the fatigability formula re-used with
    (FR -> max_rec/fat_1, maxFat -> max_rec), max_rec = maxFat/2.53 -- pmf:66-78."""
    max_rec = max_fatigue_rate / 2.53
    rec_range = max_rec / fat[0]
    return nominal_fatigabilities(n, rec_range, max_rec, ptf)


def force_to_motor_unit_count(max_force: float, conversion_factor: float) -> int:
    """Fuglevand-93 force/unit-count relation solved for the count -- muscle:189-243."""
    r = max_force / conversion_factor
    inner = np.power((1 - r) / (np.exp(4.6) - r), 1 / 4.6)
    return int(np.ceil(1 / np.log(inner)))


#: c = 1 - exp(-2 * 0.4^3), the slope constant of the linear branch -- fibers:241
LINEAR_THRESHOLD = 0.4
LINEAR_SCALE = 1 - np.exp(-2 * (0.4 ** 3))


@dataclass
class UnitTables:
    """All per-unit float64 tables of one muscle type."""
    n: int
    thr: np.ndarray
    peak: np.ndarray
    ptf: np.ndarray
    ct: np.ndarray
    fat: np.ndarray
    rec: Optional[np.ndarray]
    max_excitation: float

    @staticmethod
    def build(n: int, pool: PoolHyper = PoolHyper(), fib: FibersHyper = FibersHyper()) -> "UnitTables":
        thr = recruitment_thresholds(n, pool.max_recruitment_threshold)
        peak = peak_firing_rates(pool.max_firing_rate_first_unit, pool.max_firing_rate_last_unit,
                                 pool.max_recruitment_threshold, thr)
        ptf = peak_twitch_forces(n, fib.max_twitch_amplitude)
        ct = contraction_times(fib.max_twitch_amplitude, fib.max_contraction_time,
                               fib.contraction_time_range, ptf)
        fat = nominal_fatigabilities(n, fib.fatigability_range, fib.max_fatigue_rate, ptf)
        rec = recovery_rates(n, fib.max_fatigue_rate, fat, ptf) if fib.recovery else None
        return UnitTables(n, thr, peak, ptf, ct, fat, rec, max_excitation(pool))


# --------------------------------------------------------------------------
# One step, as pure functions on [..., n] arrays
# --------------------------------------------------------------------------

def firing_rates(exc: np.ndarray, t: UnitTables, h: PoolHyper) -> np.ndarray:
    """Un-adapted rates; its `< min_firing_rate` test IS the recruitment mask.
    pool:150-179: r = (e - thr) + minR; r[r < minR] = 0; r *= gain; r = min(r, peak)."""
    r = exc - t.thr
    r = r + h.min_firing_rate
    r = np.where(r < h.min_firing_rate, 0.0, r)
    r = r * h.firing_gain
    return np.where(r > t.peak, t.peak, r)


def recruitment_mask(exc: np.ndarray, t: UnitTables, h: PoolHyper) -> np.ndarray:
    """True where the unit fires: not ((e - thr) + minR < minR) -- pool:169-172."""
    return ~(((exc - t.thr) + h.min_firing_rate) < h.min_firing_rate)


def adaptations(r: np.ndarray, dur: np.ndarray, t: UnitTables, h: PoolHyper) -> np.ndarray:
    """Central fatigue, Eq. 12-13 -- pool:208-233.  `dur` is the duration
    BEFORE this step's update."""
    ratios = (t.thr - 1) / (h.max_recruitment_threshold - 1)                       # pool:231
    curve = h.adaptation_magnitude * (r - h.min_firing_rate + h.derecruitment_delta) * ratios  # pool:232
    exponent = -1 * (dur / h.adaptation_time_constant)                            # pool:217
    scale = 1 - np.exp(exponent)                                                  # pool:218
    a = curve * scale                                                             # pool:219
    return np.where(a < 0, 0.0, a)                                                # pool:221


def updated_durations(r: np.ndarray, dur: np.ndarray, dt, h: PoolHyper) -> np.ndarray:
    """dur[r > 0] += dt, clamped to [0, max_duration] -- pool:181-206 (r un-adapted)."""
    if not h.apply_fatigue:                                                       # pool:193-194
        return dur
    d = np.where(r > 0, dur + dt, dur)
    d = np.where(d > h.max_duration, h.max_duration, d)
    return np.where(d < 0, 0.0, d)


def current_contraction_times(cpf: np.ndarray, t: UnitTables, f: FibersHyper) -> np.ndarray:
    """ct_cur = ct * (1 + ccr * (1 - cpf/ptf)), Eq. 11 -- fibers:117-125.
    A pure function of the capacity, so it is recomputed instead of stored."""
    loss = 1 - (cpf / t.ptf)
    return t.ct * (1 + f.contraction_time_change_ratio * loss)


def normalized_forces(x: np.ndarray) -> np.ndarray:
    """Linear up to 0.4, sigmoid above -- fibers:222-247."""
    with np.errstate(over="ignore", invalid="ignore"):
        lin = (x / LINEAR_THRESHOLD) * LINEAR_SCALE                               # fibers:240-241
        sig = 1 - np.exp(-2 * np.power(x, 3))                                     # fibers:242-246
    return np.where(x <= LINEAR_THRESHOLD, lin, sig)


def fatigued_capacities(cpf: np.ndarray, nf: np.ndarray, dt, t: UnitTables, f: FibersHyper) -> np.ndarray:
    """Peripheral fatigue (+ recovery for PyMuscleFibers).
    Potvin: fibers:96-115.  PyMuscle: pmf:80-141 (strategy 4)."""
    c = cpf - (t.fat * nf) * dt                                                   # fibers:110-111
    if f.recovery:
        ratio = (t.ptf - c) / t.ptf                                               # pmf:137-139
        c = np.where(nf <= 0, c + (t.rec * ratio) * dt, c)                        # pmf:121,140-141
    c = np.where(c < 0, 0.0, c)                                                   # fibers:114 / pmf:101
    if f.recovery:
        c = np.where(c > t.ptf, t.ptf, c)                                         # pmf:104-105
    return c


# --------------------------------------------------------------------------
# Batched simulator state
# --------------------------------------------------------------------------

@dataclass
class StepRecord:
    """Everything one step produces (all [..., n] except total [...])."""
    mask: np.ndarray
    rates: np.ndarray           # adapted firing rates (Pool.step's return value)
    forces: np.ndarray          # Muscle.current_forces
    total: np.ndarray           # Fibers.step's return value


class OracleMuscles:
    """M independent, identically-parameterised muscles advanced in lock-step.

    kind='potvin'   -> PotvinFuglevandMuscle(n, central, peripheral)      muscle:95-119
    kind='standard' -> StandardMuscle(...): PyMuscleFibers, 0..1 I/O      muscle:122-279
    """

    def __init__(self, n: int, m: int = 1, kind: str = "potvin",
                 apply_central_fatigue: Optional[bool] = None,
                 apply_peripheral_fatigue: bool = True,
                 pool: Optional[PoolHyper] = None, fib: Optional[FibersHyper] = None):
        assert kind in ("potvin", "standard")
        if apply_central_fatigue is None:
            apply_central_fatigue = (kind == "potvin")                            # muscle:104 / :156
        self.kind = kind
        self.pool = pool or PoolHyper(apply_fatigue=apply_central_fatigue)
        self.fib = fib or FibersHyper(apply_fatigue=apply_peripheral_fatigue,
                                      recovery=(kind == "standard"))
        self.t = UnitTables.build(n, self.pool, self.fib)
        self.n, self.m = n, m
        self.dur = np.zeros((m, n))                                               # pool:79
        self.cpf = np.tile(self.t.ptf, (m, 1))                                    # fibers:63
        self.current_forces = np.zeros((m, n))                                    # fibers:90
        # muscle:187 -- built-in sum, i.e. sequential left-to-right
        self.max_arb_output = sum(self.t.ptf) if kind == "standard" else None

    @staticmethod
    def standard(max_force: float = 32.0, force_conversion_factor: float = 0.0123, m: int = 1,
                 apply_central_fatigue: bool = False, apply_peripheral_fatigue: bool = True):
        n = force_to_motor_unit_count(max_force, force_conversion_factor)
        return OracleMuscles(n, m, "standard", apply_central_fatigue, apply_peripheral_fatigue)

    # -- the two halves, usable on their own (Pool.step / Fibers.step) -----
    def pool_step(self, exc: np.ndarray, dt) -> tuple[np.ndarray, np.ndarray]:
        """pool:105-126.  exc is [m, n].  Returns (mask, adapted rates)."""
        r = firing_rates(exc, self.t, self.pool)
        a = adaptations(r, self.dur, self.t, self.pool)
        adapted = r - a                                                           # pool:121
        self.dur = updated_durations(r, self.dur, dt, self.pool)                  # pool:124
        return recruitment_mask(exc, self.t, self.pool), adapted

    def fibers_step(self, rates: np.ndarray, dt) -> tuple[np.ndarray, np.ndarray]:
        """fibers:261-282.  Returns (per-unit forces, per-muscle total)."""
        ct_cur = current_contraction_times(self.cpf, self.t, self.fib)
        x = ct_cur * (rates / 1000)                                               # fibers:220
        nf = normalized_forces(x)
        forces = nf * self.cpf                                                    # fibers:258
        total = np.sum(forces, axis=-1)                                           # fibers:276
        if self.fib.apply_fatigue:                                                # fibers:279-280
            self.cpf = fatigued_capacities(self.cpf, nf, dt, self.t, self.fib)
        self.current_forces = forces
        return forces, total

    def step(self, exc, dt) -> StepRecord:
        """Muscle.step muscle:65-92 (+ StandardMuscle.step muscle:258-279).

        exc: scalar, [m] (one value per muscle) or [m, n] (per unit)."""
        e = np.asarray(exc, dtype=np.float64)
        if e.ndim <= 1:
            e = np.broadcast_to(e.reshape(-1, 1) if e.ndim == 1 else e, (self.m, self.n))
        if self.kind == "standard":
            e = e * self.t.max_excitation                                         # muscle:275
        mask, rates = self.pool_step(e, dt)
        forces, total = self.fibers_step(rates, dt)
        if self.kind == "standard":
            total = total / self.max_arb_output                                   # muscle:278
        return StepRecord(mask, rates, forces, total)

    def peripheral_fatigue(self) -> np.ndarray:
        """StandardMuscle.get_peripheral_fatigue muscle:248-256 (built-in sum)."""
        out = np.empty(self.m)
        for i in range(self.m):
            out[i] = 1 - sum(self.cpf[i]) / self.max_arb_output
        return out


# --------------------------------------------------------------------------
# Reference-shaped single muscle (one object per muscle, [n] arrays): this is
# the form the cpu_baseline times, because the reference's callers loop
# `for muscle in muscles: muscle.step(e, dt)` (examples/envs/muscled_hopper.py:33-34).
# --------------------------------------------------------------------------

class OracleMuscle:
    """One muscle; same per-step NumPy dispatch pattern as the reference's
    Muscle -> Pool -> Fibers object graph (about 70 small-array ufunc calls)."""

    def __init__(self, n: int, kind: str = "potvin", **kw):
        self._b = OracleMuscles(n, 1, kind, **kw)
        self._t = self._b.t
        self._dur = np.zeros(n)
        self._cpf = self._t.ptf.copy()
        self._ct_cur = self._t.ct.copy()
        self.current_forces = np.zeros(n)
        self.motor_unit_count = n

    def step(self, exc, dt):
        b, t, ph, fh = self._b, self._t, self._b.pool, self._b.fib
        if isinstance(exc, (float, int)):                                         # muscle:81-86
            exc = np.full(t.n, exc)
        e = np.array(exc)                                                         # muscle:89
        if b.kind == "standard":
            e = e * t.max_excitation
        # pool:150-179
        r = e - t.thr
        r += ph.min_firing_rate
        r[r < ph.min_firing_rate] = 0
        r *= ph.firing_gain
        over = r > t.peak
        r[over] = t.peak[over]
        # pool:208-233
        ratios = (t.thr - 1) / (ph.max_recruitment_threshold - 1)
        a = ph.adaptation_magnitude * (r - ph.min_firing_rate + ph.derecruitment_delta) * ratios
        a = a * (1 - np.exp(-1 * (self._dur / ph.adaptation_time_constant)))
        a[a < 0] = 0.0
        fr = r - a
        # pool:181-206
        if ph.apply_fatigue:
            self._dur[r > 0] += dt
            self._dur[self._dur > ph.max_duration] = ph.max_duration
            self._dur[self._dur < 0] = 0
        # fibers:211-259
        x = self._ct_cur * (fr / 1000)
        nf = x.copy()
        lo = nf <= LINEAR_THRESHOLD
        hi = nf > LINEAR_THRESHOLD
        nf[lo] /= 0.4
        nf[lo] *= LINEAR_SCALE
        nf[hi] = 1 - np.exp(-2 * np.power(nf[hi], 3))
        self.current_forces = nf * self._cpf
        total = np.sum(self.current_forces)
        # fibers:96-125 / pmf:80-141
        if fh.apply_fatigue:
            self._cpf -= (t.fat * nf) * dt
            if fh.recovery:
                rc = nf <= 0
                peak = t.ptf[rc]
                self._cpf[rc] += (t.rec[rc] * ((peak - self._cpf[rc]) / peak)) * dt
            self._cpf[self._cpf < 0] = 0.0
            if fh.recovery:
                ov = self._cpf > t.ptf
                self._cpf[ov] = t.ptf[ov]
            self._ct_cur = t.ct * (1 + fh.contraction_time_change_ratio * (1 - (self._cpf / t.ptf)))
        if b.kind == "standard":
            return total / b.max_arb_output
        return total
