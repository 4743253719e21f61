"""
Drop-in pool and fibers models.  Same constructor arguments, attributes and
error behaviour as the reference classes; the arithmetic runs in the K1 kernel
(stage POOL / stage FIBERS of mb_step) on a batch of one, in float64.

Reference: pymuscle/potvin_fuglevand_2017_motor_neuron_pool.py (Pool),
pymuscle/potvin_fuglevand_2017_muscle_fibers.py (Fibers),
pymuscle/pymuscle_fibers.py (PyMuscleFibers).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _native as nat
from .engine import MuscleBatch
from .model import Model
from .tables import FibersParams, MuscleSpec, PoolParams, fibers_tables, pool_tables


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("pymuscle_b200 needs a CUDA device: there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


class PotvinFuglevand2017MotorNeuronPool(Model):
    """Potvin & Fuglevand 2017 motor-neuron pool (pool:8-295).

    ``step(excitations[n], step_size) -> firing_rates[n]`` with adaptation
    (central fatigue) and on-duration bookkeeping, evaluated on the GPU.
    """

    def __init__(self, motor_unit_count: int, max_recruitment_threshold: int = 50, firing_gain: float = 1.0,
                 min_firing_rate: int = 8, max_firing_rate_first_unit: int = 35,
                 max_firing_rate_last_unit: int = 25, derecruitment_delta: int = 2,
                 adaptation_magnitude: float = 0.67, adaptation_time_constant: float = 22.0,
                 max_duration: float = 20000.0, apply_fatigue: bool = True):
        super().__init__(motor_unit_count)
        self._params = PoolParams(max_recruitment_threshold, firing_gain, min_firing_rate,
                                  max_firing_rate_first_unit, max_firing_rate_last_unit, derecruitment_delta,
                                  adaptation_magnitude, adaptation_time_constant, max_duration, apply_fatigue)
        self._recruitment_thresholds, self._peak_firing_rates = pool_tables(motor_unit_count, self._params)
        self._max_recruitment_threshold = max_recruitment_threshold
        self._firing_gain = firing_gain
        self._min_firing_rate = min_firing_rate
        self._derecruitment_delta = derecruitment_delta
        self._adaptation_magnitude = adaptation_magnitude
        self._adaptation_time_constant = adaptation_time_constant
        self._max_duration = max_duration
        self._apply_fatigue = apply_fatigue
        self.max_excitation = self._params.max_excitation                      # pool:96-99
        self._engine = MuscleBatch(MuscleSpec(motor_unit_count, self._params, FibersParams()), 1,
                                   _device(), "f64")

    @property
    def _recruitment_durations(self) -> np.ndarray:
        return self._engine.dur[0].cpu().numpy()

    def step(self, motor_pool_input, step_size: float) -> np.ndarray:
        """pool:283-295.  A scalar raises TypeError (len of unsized object), a wrong
        length AssertionError, exactly as the reference."""
        assert (len(motor_pool_input) == self.motor_unit_count)
        x = np.ascontiguousarray(motor_pool_input, dtype=np.float64)
        rates = torch.empty(1, self.motor_unit_count, dtype=torch.float64, device=self._engine.device)
        self._engine.step(x, step_size, stage=nat.MB_STAGE_POOL, rates_out=rates)
        return rates[0].cpu().numpy()


class PotvinFuglevand2017MuscleFibers(Model):
    """Potvin & Fuglevand 2017 muscle fibers (fibers:9-300): firing rates ->
    per-unit forces and their total, with peripheral fatigue."""

    _recovery = False

    def __init__(self, motor_unit_count: int, max_twitch_amplitude: int = 100, max_contraction_time: int = 90,
                 contraction_time_range: int = 3, max_fatigue_rate: float = 0.0225, fatigability_range: int = 180,
                 contraction_time_change_ratio: float = 0.379, apply_fatigue: bool = True):
        super().__init__(motor_unit_count)
        self._params = FibersParams(max_twitch_amplitude, max_contraction_time, contraction_time_range,
                                    max_fatigue_rate, fatigability_range, contraction_time_change_ratio,
                                    apply_fatigue, recovery=self._recovery)
        ptf, ct, fat, rec = fibers_tables(motor_unit_count, self._params)
        self._peak_twitch_forces = ptf
        self._contraction_times = ct
        self._nominal_fatigabilities = fat
        if rec is not None:
            self._recovery_rates = rec
        self._contraction_time_change_ratio = contraction_time_change_ratio
        self._apply_fatigue = apply_fatigue
        self._max_fatigue_rate = max_fatigue_rate
        self._engine = MuscleBatch(MuscleSpec(motor_unit_count, PoolParams(), self._params), 1, _device(), "f64")
        self._forces_host = np.zeros(motor_unit_count)                         # fibers:90

    @property
    def current_forces(self) -> np.ndarray:
        """Per-unit forces of the last step: a new array per step (fibers:258)."""
        if self._forces_host is None:
            src = getattr(self, "_forces_src", None)            # set by Muscle.step (fused launch, host I/O)
            self._forces_host = src.host_forces()[0] if src is not None else \
                self._engine.current_forces[0].cpu().numpy()
        return self._forces_host

    def __getstate__(self):
        d = dict(self.__dict__)
        d["_forces_host"] = np.array(self.current_forces)      # materialise: the source engine does not travel
        d["_forces_src"] = None
        return d

    @property
    def current_peak_forces(self) -> np.ndarray:
        """fibers:92-94 (a host snapshot of the device state)."""
        return self._engine.cpf[0].cpu().numpy()

    @property
    def _current_peak_forces(self) -> np.ndarray:
        return self.current_peak_forces

    def step(self, motor_pool_output, step_size: float) -> float:
        """fibers:284-300"""
        assert (len(motor_pool_output) == self.motor_unit_count)
        x = np.ascontiguousarray(motor_pool_output, dtype=np.float64)
        total = self._engine.step(x, step_size, stage=nat.MB_STAGE_FIBERS)
        self._forces_src = None
        self._forces_host = None
        return np.float64(total.item())


class PyMuscleFibers(PotvinFuglevand2017MuscleFibers):
    """PyMuscle's standard fibers: the Potvin fibers plus recovery of resting
    units and an upper clamp at the peak twitch force (pmf:5-141)."""

    _recovery = True

    def __init__(self, *args, force_conversion_factor: float = 0.028, **kwargs):
        super().__init__(*args, **kwargs)
        self.force_conversion_factor = force_conversion_factor                 # pmf:63-64
