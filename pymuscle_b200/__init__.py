"""
pymuscle_b200 -- the PyMuscle per-step hot path (Potvin-Fuglevand 2017 motor
neuron pool -> muscle fibers force / fatigue -> per-muscle total) as
hand-written sm_100a CUDA, behind the reference's own Python surface
(pymuscle/__init__.py:1-12) plus a batched entry point, MuscleBatch.
"""
from .__version__ import __version__  # noqa: F401
from .model import Model  # noqa: F401
from .muscle import Muscle  # noqa: F401
from .muscle import PotvinFuglevandMuscle  # noqa: F401
from .muscle import StandardMuscle  # noqa: F401
from .models import PotvinFuglevand2017MuscleFibers  # noqa: F401
from .models import PotvinFuglevand2017MotorNeuronPool  # noqa: F401
from .models import PyMuscleFibers  # noqa: F401
from .engine import MuscleBatch  # noqa: F401
from .tables import MuscleSpec, PoolParams, FibersParams  # noqa: F401
from .distributed import ShardedMuscleBatch, shard_bounds  # noqa: F401
from .hill_type import contractile_element_force_velocity_curve  # noqa: F401
from .hill_type import contractile_element_force_length_curve  # noqa: F401
from .envs import MuscleGroupEnv  # noqa: F401
from .history import ForceHistory  # noqa: F401
