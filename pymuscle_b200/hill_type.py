"""
Hill-type helper curves (reference: pymuscle/hill_type.py:15-66).

Two dimensionless multipliers a physics coupling applies to a muscle's active force: the
force-length relation (a Gaussian-like bump around the optimal length) and the
force-velocity relation (a sigmoid between 0 and the eccentric maximum).  They are not on
the per-step path of the reference (nothing in pymuscle calls them); here they accept
floats, NumPy arrays and torch tensors so that a vector environment can evaluate them on
the device for every (agent, muscle) and hand the product to the step kernel's epilogue
(``MuscleBatch.step_env(gain=...)``).
"""
from __future__ import annotations

import numpy as np

try:                                    # torch is optional for the scalar helpers
    import torch
except Exception:                       # pragma: no cover
    torch = None


def _is_tensor(*xs):
    return torch is not None and any(isinstance(x, torch.Tensor) for x in xs)


def contractile_element_force_length_curve(rest_length, current_length, curve_width_factor: float = 17.33,
                                           peak_force_length: float = 1.1):
    """Fraction of the maximal contractile force available at ``current_length``
    (hill_type.py:15-41).  As in the reference, the optimal length is fixed at 1.10 rest
    lengths (Aubert 1951) whatever ``peak_force_length`` says (hill_type.py:38)."""
    optimum = 1.10
    norm_length = current_length / rest_length
    if _is_tensor(rest_length, current_length):
        return torch.exp(-curve_width_factor * torch.abs(norm_length - optimum) ** 3)
    return np.exp(curve_width_factor * (-1 * (abs(norm_length - optimum) ** 3)))


def contractile_element_force_velocity_curve(rest_length, current_length, prev_length, time_step: float,
                                             max_velocity: float = 3.0, max_exccentric_multiple: float = 1.8):
    """Force multiplier in (0, max_exccentric_multiple) from the shortening velocity over the last
    step (hill_type.py:44-66)."""
    velocity = ((prev_length - current_length) / rest_length) / time_step
    exponent = (0.04 - velocity / max_velocity) / 0.18
    e = torch.exp(exponent) if _is_tensor(rest_length, current_length, prev_length) else np.exp(exponent)
    return max_exccentric_multiple - (max_exccentric_multiple / (1 + e))
