"""The Model protocol Muscle composes (reference: pymuscle/model.py:4-18)."""


class Model(object):
    """Base class of pool / fibers models: stores the unit count, ``step`` is abstract."""

    def __init__(self, motor_unit_count: int):
        self.motor_unit_count = motor_unit_count

    def step(self, inputs, step_size: float):
        raise NotImplementedError
