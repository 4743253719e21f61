"""
pymuscle_b200: B200-native batched motor-unit simulator, API-compatible with
PyMuscle 0.1.2 on the per-step hot path.
"""
VERSION = (0, 1, 2)

__version__ = '.'.join(map(str, VERSION))
