"""
Host-side parameter tables and muscle specifications.

The per-unit float64 tables are built with NumPy using the reference's own
expressions and operation order, because the recruitment mask is a bit-exact
contract and the thresholds contain ``exp``/``log`` whose last bit depends on
the library that evaluates them.  They are built
once per muscle type, packed by ``mb_params_pack`` (C ABI) into the layout the
kernels load, and uploaded.

Reference: pymuscle/potvin_fuglevand_2017_motor_neuron_pool.py:235-281 (pool
tables), pymuscle/potvin_fuglevand_2017_muscle_fibers.py:127-209 (fibers
tables), pymuscle/pymuscle_fibers.py:66-78 (recovery rates),
pymuscle/muscle.py:189-243 (unit count from force).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field, replace
from typing import Optional, Sequence

import numpy as np

from . import _native as nat


@dataclass(frozen=True)
class PoolParams:
    """Constructor arguments of PotvinFuglevand2017MotorNeuronPool (pool:53-66)."""
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

    @property
    def max_excitation(self):
        """pool:96-99"""
        return self.max_recruitment_threshold + \
            (self.max_firing_rate_last_unit - self.min_firing_rate) / self.firing_gain


@dataclass(frozen=True)
class FibersParams:
    """Constructor arguments of the fibers models (fibers:46-56, pmf:55-60)."""
    max_twitch_amplitude: float = 100
    max_contraction_time: float = 90
    contraction_time_range: float = 3
    max_fatigue_rate: float = 0.0225
    fatigability_range: float = 180
    contraction_time_change_ratio: float = 0.379
    apply_fatigue: bool = True
    recovery: bool = False            # True = PyMuscleFibers


def _indices(n: int) -> np.ndarray:
    return np.arange(1, n + 1)


def pool_tables(n: int, p: PoolParams):
    """(thresholds, peak firing rates) -- pool:263-281, pool:235-261."""
    i = _indices(n)
    thr = np.exp((np.log(p.max_recruitment_threshold) * (i - 1)) / (n - 1))
    span = p.max_firing_rate_first_unit - p.max_firing_rate_last_unit
    peak = p.max_firing_rate_first_unit - (span * ((thr - thr[0]) / (p.max_recruitment_threshold - thr[0])))
    return thr, peak


def _fatigabilities(n, rng, max_rate, ptf):
    """fibers:181-209 (also re-used for the recovery rates, pmf:73-78)."""
    i = _indices(n)
    curve = np.exp((np.log(rng) / (n - 1)) * (i - 1))
    return curve * (max_rate / rng) * ptf


def fibers_tables(n: int, f: FibersParams):
    """(peak twitch forces, contraction times, fatigabilities, recovery rates or None)
    -- fibers:163-179, fibers:127-161, fibers:181-209, pmf:66-78."""
    i = _indices(n)
    ptf = np.exp((np.log(f.max_twitch_amplitude) * (i - 1)) / (n - 1))
    scale = np.log(f.max_twitch_amplitude) / np.log(f.contraction_time_range)
    ct = f.max_contraction_time * np.power(1 / ptf, 1 / scale)
    fat = _fatigabilities(n, f.fatigability_range, f.max_fatigue_rate, ptf)
    rec = None
    if f.recovery:
        max_rec = f.max_fatigue_rate / 2.53
        rec = _fatigabilities(n, max_rec / fat[0], max_rec, ptf)
    return ptf, ct, fat, rec


def force_to_motor_unit_count(max_force: float, conversion_factor: float) -> int:
    """muscle.py:189-243"""
    r = max_force / conversion_factor
    inner = np.power((1 - r) / (np.exp(4.6) - r), 1 / 4.6)
    return int(np.ceil(1 / np.log(inner)))


@dataclass(frozen=True)
class MuscleSpec:
    """One muscle type: unit count, model hyper-parameters and I/O scaling."""
    n: int
    pool: PoolParams = PoolParams()
    fibers: FibersParams = FibersParams()
    input_scale: float = 1.0          # StandardMuscle: pool.max_excitation (muscle.py:275)
    output_divisor: float = 1.0       # StandardMuscle: max_arb_output      (muscle.py:278)
    name: str = "muscle"

    def __post_init__(self):
        if int(self.n) < 2:
            raise ValueError("a muscle needs at least 2 motor units (the reference divides by n-1)")

    @staticmethod
    def potvin(motor_unit_count: int, apply_central_fatigue: bool = True,
               apply_peripheral_fatigue: bool = True) -> "MuscleSpec":
        """PotvinFuglevandMuscle (muscle.py:95-119)."""
        return MuscleSpec(int(motor_unit_count), PoolParams(apply_fatigue=apply_central_fatigue),
                          FibersParams(apply_fatigue=apply_peripheral_fatigue), name="potvin")

    @staticmethod
    def standard(max_force: float = 32.0, force_conversion_factor: float = 0.0123,
                 apply_central_fatigue: bool = False, apply_peripheral_fatigue: bool = True) -> "MuscleSpec":
        """StandardMuscle (muscle.py:122-187): PyMuscleFibers, 0..1 input and output."""
        n = force_to_motor_unit_count(max_force, force_conversion_factor)
        pool = PoolParams(apply_fatigue=apply_central_fatigue)
        fib = FibersParams(apply_fatigue=apply_peripheral_fatigue, recovery=True)
        ptf = fibers_tables(n, fib)[0]
        return MuscleSpec(n, pool, fib, input_scale=pool.max_excitation,
                          output_divisor=sum(ptf),              # muscle.py:187 (built-in, sequential sum)
                          name="standard")

    def tables(self) -> dict:
        thr, peak = pool_tables(self.n, self.pool)
        ptf, ct, fat, rec = fibers_tables(self.n, self.fibers)
        return dict(thr=thr, peak=peak, ptf=ptf, ct=ct, fat=fat, rec=rec)


def _shared_scalars(spec: MuscleSpec):
    """The scalars one launch shares (everything in MbConfig except the divisor)."""
    p, f = spec.pool, spec.fibers
    return (p.max_recruitment_threshold, p.firing_gain, p.min_firing_rate, p.derecruitment_delta,
            p.adaptation_magnitude, p.adaptation_time_constant, p.max_duration, p.apply_fatigue,
            f.contraction_time_change_ratio, f.apply_fatigue, f.recovery, spec.input_scale)


@dataclass
class PackedRow:
    """The packed parameter table of one batch row (S muscles back to back)."""
    specs: Sequence[MuscleSpec]
    precision: int
    cfg: nat.MbConfig
    blob: np.ndarray                  # uint8, mb_params_bytes long
    seg_off: np.ndarray               # int32 [S+1]
    seg_div: np.ndarray               # float64 [S]
    ptf: np.ndarray                   # float64 [L] initial capacities (fibers:63)

    @property
    def L(self) -> int:
        return int(self.seg_off[-1])

    @property
    def S(self) -> int:
        return len(self.specs)


def make_config(spec: MuscleSpec, precision: int) -> nat.MbConfig:
    p, f = spec.pool, spec.fibers
    cfg = nat.MbConfig()
    cfg.precision = precision
    cfg.fibers_kind = nat.MB_FIBERS_PYMUSCLE if f.recovery else nat.MB_FIBERS_POTVIN
    cfg.central_fatigue = int(bool(p.apply_fatigue))
    cfg.peripheral_fatigue = int(bool(f.apply_fatigue))
    cfg.max_recruitment_threshold = float(p.max_recruitment_threshold)
    cfg.firing_gain = float(p.firing_gain)
    cfg.min_firing_rate = float(p.min_firing_rate)
    cfg.derecruitment_delta = float(p.derecruitment_delta)
    cfg.adaptation_magnitude = float(p.adaptation_magnitude)
    cfg.adaptation_time_constant = float(p.adaptation_time_constant)
    cfg.max_duration = float(p.max_duration)
    cfg.contraction_time_change_ratio = float(f.contraction_time_change_ratio)
    cfg.input_scale = float(spec.input_scale)
    cfg.output_divisor = float(spec.output_divisor)
    cfg.table_props = 0
    return cfg


def pack_row(specs: Sequence[MuscleSpec], precision: int) -> PackedRow:
    """Concatenate the tables of a row's muscles and pack them through the C ABI."""
    specs = list(specs)
    if not specs:
        raise ValueError("at least one MuscleSpec is required")
    first = _shared_scalars(specs[0])
    for s in specs[1:]:
        if _shared_scalars(s) != first:
            raise ValueError("all muscles of a batch row must share their scalar hyper-parameters "
                             "(they may differ in unit count and output divisor)")
    tabs = [s.tables() for s in specs]
    cat = {k: np.ascontiguousarray(np.concatenate([t[k] for t in tabs]), dtype=np.float64)
           for k in ("thr", "peak", "ptf", "ct", "fat")}
    rec = None
    if specs[0].fibers.recovery:
        rec = np.ascontiguousarray(np.concatenate([t["rec"] for t in tabs]), dtype=np.float64)
    seg_off = np.zeros(len(specs) + 1, dtype=np.int32)
    seg_off[1:] = np.cumsum([s.n for s in specs])
    L = int(seg_off[-1])
    cfg = make_config(specs[0], precision)
    lib = nat.lib()
    nbytes = lib.mb_params_bytes(C.byref(cfg), L)
    blob = np.zeros(nbytes, dtype=np.uint8)
    props = C.c_int32(0)

    def ptr(a):
        return a.ctypes.data_as(C.c_void_p) if a is not None else None

    nat.check(lib.mb_params_pack(C.byref(cfg), L, ptr(cat["thr"]), ptr(cat["peak"]), ptr(cat["ptf"]),
                                 ptr(cat["ct"]), ptr(cat["fat"]), ptr(rec), ptr(blob), C.byref(props)))
    cfg.table_props = props.value
    seg_div = np.array([float(s.output_divisor) for s in specs], dtype=np.float64)
    return PackedRow(specs, precision, cfg, blob, seg_off, seg_div, cat["ptf"])


def pack_rows_per_instance(row_specs: Sequence[MuscleSpec], precision: int) -> PackedRow:
    """Per-instance hyper-parameters (domain randomisation): one packed table per batch row.

    Every row is one muscle of the same unit count; rows may differ in every hyper-parameter
    that only enters the per-unit tables -- recruitment range, peak firing rates, derecruitment
    delta, adaptation magnitude, twitch amplitudes, contraction times, fatigue / recovery rates,
    contraction-time change ratio -- but must share the scalars a launch passes by value
    (firing gain, minimal firing rate, adaptation time constant, maximal duration, the fatigue
    switches, the fibers model and the input / output scaling).  Returns a PackedRow whose blob
    holds ``len(row_specs)`` consecutive tables and whose cfg.param_rows says so."""
    row_specs = list(row_specs)
    if len(row_specs) < 2:
        raise ValueError("per-instance tables need at least two rows")
    first = row_specs[0]

    def launch_scalars(s: MuscleSpec):
        p, f = s.pool, s.fibers
        return (s.n, p.firing_gain, p.min_firing_rate, p.adaptation_time_constant, p.max_duration, p.apply_fatigue,
                f.apply_fatigue, f.recovery, s.input_scale, s.output_divisor)

    for r, s in enumerate(row_specs[1:], 1):
        if launch_scalars(s) != launch_scalars(first):
            raise ValueError(f"row {r} differs from row 0 in a per-launch scalar (unit count, firing gain, minimal "
                             "firing rate, adaptation time constant, maximal duration, switches or I/O scaling)")
    packs = [pack_row([s], precision) for s in row_specs]
    cfg = packs[0].cfg
    props = cfg.table_props
    for pk in packs[1:]:
        props &= pk.cfg.table_props                       # a fast path must hold for every row
    cfg.table_props = props
    cfg.param_rows = len(row_specs)
    blob = np.concatenate([pk.blob for pk in packs])
    ptf = np.stack([pk.ptf for pk in packs])               # [rows, n] initial capacities
    return PackedRow([first], precision, cfg, blob, packs[0].seg_off, packs[0].seg_div, ptf)
