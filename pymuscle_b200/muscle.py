"""
Drop-in Muscle façade (reference: pymuscle/muscle.py).

``Muscle(pool, fibers)`` composes any two ``Model``s as the reference does.
When both are this package's native models the step is ONE fused K1 launch
(pool -> fibers -> total) over the models' own device state; otherwise the two
``step`` methods are chained generically.
"""
from __future__ import annotations

from typing import Union

import numpy as np
import torch

from . import _native as nat
from .engine import MuscleBatch
from .model import Model
from .models import (PotvinFuglevand2017MotorNeuronPool, PotvinFuglevand2017MuscleFibers, PyMuscleFibers)
from .tables import MuscleSpec, force_to_motor_unit_count


class Muscle(object):
    """Excitation in, total fibre force out, one call per time step (muscle.py:14-92)."""

    def __init__(self, motor_neuron_pool_model: Model, muscle_fibers_model: Model):
        assert motor_neuron_pool_model.motor_unit_count == muscle_fibers_model.motor_unit_count
        self._pool = motor_neuron_pool_model
        self._fibers = muscle_fibers_model
        self._engine = None
        self._input_scale = 1.0
        self._output_divisor = 1.0
        if isinstance(self._pool, PotvinFuglevand2017MotorNeuronPool) and \
                isinstance(self._fibers, PotvinFuglevand2017MuscleFibers):
            self._build_engine()

    def _build_engine(self):
        """One fused engine that shares the two models' device state tensors."""
        spec = MuscleSpec(self._pool.motor_unit_count, self._pool._params, self._fibers._params,
                          input_scale=self._input_scale, output_divisor=self._output_divisor)
        eng = MuscleBatch(spec, 1, self._pool._engine.device, "f64")
        eng.dur = self._pool._engine.dur
        eng.cpf = self._fibers._engine.cpf
        self._engine = eng

    # copying / pickling: the fused engine aliases the two models' device state, so it is rebuilt
    def __getstate__(self):
        d = dict(self.__dict__)
        d["_engine"] = d["_engine"] is not None
        return d

    def __setstate__(self, d):
        had_engine = d.pop("_engine")
        self.__dict__.update(d)
        self._engine = None
        if had_engine:
            self._build_engine()
            # the last step's forces travelled with the old engine's host buffer: keep a copy
            if getattr(self._fibers, "_forces_src", None) is not None:
                self._fibers._forces_src = None

    @property
    def motor_unit_count(self):
        return self._pool.motor_unit_count

    @property
    def max_excitation(self):
        return self._pool.max_excitation

    @property
    def current_forces(self):
        return self._fibers.current_forces

    def step(self, motor_pool_input: Union[int, float, np.ndarray], step_size: float) -> float:
        """muscle.py:65-92"""
        if self._engine is None:
            if isinstance(motor_pool_input, float) or isinstance(motor_pool_input, int):
                motor_pool_input = np.full(self._pool.motor_unit_count, motor_pool_input)
            return self._fibers.step(self._pool.step(np.array(motor_pool_input), step_size), step_size)
        if isinstance(motor_pool_input, float) or isinstance(motor_pool_input, int):
            x = float(motor_pool_input)                                        # muscle.py:81-86
        else:
            x = np.array(motor_pool_input, dtype=np.float64)                   # muscle.py:89
            assert (len(x) == self.motor_unit_count)                           # pool:294
        # one launch + one stream sync: excitation, total and per-unit forces go through pinned
        # host memory that the kernel reads / writes directly (MuscleBatch.step_hostio)
        total = self._engine.step_hostio(x, step_size)[0, 0]
        # the fibers model hands out the per-unit forces of this step as a fresh array (fibers:258)
        self._fibers._forces_src = self._engine
        self._fibers._forces_host = None
        return np.float64(total)


class PotvinFuglevandMuscle(Muscle):
    """Potvin pool + Potvin fibers (muscle.py:95-119)."""

    def __init__(self, motor_unit_count: int, apply_central_fatigue: bool = True,
                 apply_peripheral_fatigue: bool = True):
        pool = PotvinFuglevand2017MotorNeuronPool(motor_unit_count, apply_fatigue=apply_central_fatigue)
        fibers = PotvinFuglevand2017MuscleFibers(motor_unit_count, apply_fatigue=apply_peripheral_fatigue)
        super().__init__(motor_neuron_pool_model=pool, muscle_fibers_model=fibers)


class StandardMuscle(Muscle):
    """Potvin pool + PyMuscleFibers with inputs and outputs in 0..1 (muscle.py:122-279)."""

    def __init__(self, max_force: float = 32.0, force_conversion_factor: float = 0.0123,
                 apply_central_fatigue: bool = False, apply_peripheral_fatigue: bool = True):
        self.max_force = max_force
        self.force_conversion_factor = force_conversion_factor
        motor_unit_count = self.force_to_motor_unit_count(self.max_force, self.force_conversion_factor)
        pool = PotvinFuglevand2017MotorNeuronPool(motor_unit_count, apply_fatigue=apply_central_fatigue)
        fibers = PyMuscleFibers(motor_unit_count, force_conversion_factor=force_conversion_factor,
                                apply_fatigue=apply_peripheral_fatigue)
        super().__init__(motor_neuron_pool_model=pool, muscle_fibers_model=fibers)
        self.max_arb_output = sum(self._fibers._peak_twitch_forces)            # muscle.py:187
        # fold the 0..1 scaling into the fused launch (muscle.py:275, :278)
        self._input_scale = self.max_excitation
        self._output_divisor = self.max_arb_output
        self._build_engine()

    force_to_motor_unit_count = staticmethod(force_to_motor_unit_count)

    def get_central_fatigue(self):
        raise NotImplementedError

    def get_peripheral_fatigue(self) -> float:
        """0 = rested ... 1 = completely fatigued (muscle.py:248-256)."""
        return np.float64(self._engine.get_peripheral_fatigue().item())

    def step(self, motor_pool_input: Union[int, float, np.ndarray], step_size: float) -> float:
        """muscle.py:258-279.  Unlike the reference this does NOT scale a caller's
        ndarray in place; the scaling happens in the kernel."""
        return super().step(motor_pool_input, step_size)
