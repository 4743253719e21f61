"""Hill-type helper curves against values produced by the unmodified reference
(tests/golden/hill.npz, written by tests/golden/make_golden.py)."""
import numpy as np
import torch

from conftest import golden
from pymuscle_b200 import (contractile_element_force_length_curve,
                           contractile_element_force_velocity_curve)


def test_curves_match_the_reference_bit_for_bit_on_numpy_inputs():
    g = golden("hill")
    fl = contractile_element_force_length_curve(g["rest"], g["cur"])
    fv = contractile_element_force_velocity_curve(g["rest"], g["cur"], g["prev"], float(g["dt"]))
    assert np.array_equal(fl, g["fl"]) and np.array_equal(fv, g["fv"])


def test_scalar_and_tensor_forms():
    g = golden("hill")
    assert contractile_element_force_length_curve(1.0, 1.1) == 1.0              # optimum at 1.10 rest lengths
    assert contractile_element_force_length_curve(1.0, 1.1, peak_force_length=0.9) == 1.0   # reference quirk: ignored
    t = lambda a: torch.from_numpy(a)
    fl = contractile_element_force_length_curve(t(g["rest"]), t(g["cur"]))
    fv = contractile_element_force_velocity_curve(t(g["rest"]), t(g["cur"]), t(g["prev"]), float(g["dt"]))
    # torch's exp differs from NumPy's in the last bit, and 1.8 - 1.8/(1 + e) cancels near 0
    np.testing.assert_allclose(fl.numpy(), g["fl"], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(fv.numpy(), g["fv"], rtol=1e-13, atol=1e-15)
    assert ((fv.numpy() > 0) & (fv.numpy() < 1.8)).all()
