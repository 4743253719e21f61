"""
The drop-in classes against the values and error contracts the reference's own
test-suite pins, plus the README loop (BASELINE config #1).
Written to read like the reference's tests; every value is the reference's.
"""
import re

import numpy as np
import pytest

from conftest import golden

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _need_gpu(cuda_device):
    return cuda_device


def test_module_exports_and_version():
    import pymuscle_b200 as pymuscle
    for name in ("Model", "Muscle", "PotvinFuglevandMuscle", "StandardMuscle", "PotvinFuglevand2017MuscleFibers",
                 "PotvinFuglevand2017MotorNeuronPool", "PyMuscleFibers", "MuscleBatch"):
        assert hasattr(pymuscle, name)
    assert re.match(r"\d\.\d\.\d", pymuscle.__version__)


def test_model_protocol():
    from pymuscle_b200 import Model
    with pytest.raises(TypeError):
        Model()
    m = Model(120)
    assert m.motor_unit_count == 120
    with pytest.raises(TypeError):
        m.step()
    with pytest.raises(NotImplementedError):
        m.step(20, 1)


def test_pool():
    from pymuscle_b200 import PotvinFuglevand2017MotorNeuronPool as Pool
    with pytest.raises(TypeError):
        Pool()
    p = Pool(120)
    assert p.motor_unit_count == 120 and p.max_excitation == 67.0
    with pytest.raises(TypeError):
        p.step()
    with pytest.raises(TypeError):
        p.step(33.0, 1)
    with pytest.raises(AssertionError):
        p.step(np.ones(3), 1)
    assert sum(p.step(np.zeros(120), 1.0)) == pytest.approx(0.0)
    for e, want in ((40.0, 3503.58881), (67.0, 3915.06787), (107.0, 3915.06787)):
        assert np.sum(Pool(120).step(np.full(120, e), 1.0)) == pytest.approx(want)
    p = Pool(120)
    first = np.sum(p.step(np.full(120, 67.0), 1.0))
    for _ in range(10):
        adapted = np.sum(p.step(np.full(120, 67.0), 1.0))
    assert adapted != pytest.approx(first) and adapted == pytest.approx(3749.91061)
    p = Pool(120, apply_fatigue=False)
    first = np.sum(p.step(np.full(120, 67.0), 1.0))
    for _ in range(10):
        nxt = np.sum(p.step(np.full(120, 67.0), 1.0))
    assert nxt == pytest.approx(first)


@pytest.mark.parametrize("cls_name", ["PotvinFuglevand2017MuscleFibers", "PyMuscleFibers"])
def test_fibers(cls_name):
    import pymuscle_b200
    Fibers = getattr(pymuscle_b200, cls_name)
    with pytest.raises(TypeError):
        Fibers()
    f = Fibers(120)
    assert f.motor_unit_count == 120
    assert np.equal(f.current_forces, np.zeros(120)).all()
    with pytest.raises(TypeError):
        f.step()
    with pytest.raises(TypeError):
        f.step(33.0)
    with pytest.raises(TypeError):
        f.step(33.0, 1)
    with pytest.raises(AssertionError):
        f.step(np.ones(3), 1)
    assert f.step(np.zeros(120), 1.0) == pytest.approx(0.0)
    for r, want in ((40.0, 2586.7530897), (67.0, 2609.0308816), (107.0, 2609.0308816)):
        assert Fibers(120).step(np.full(120, r), 1.0) == pytest.approx(want)


def test_pymuscle_fibers_fatigue():
    from pymuscle_b200 import PyMuscleFibers as Fibers
    f = Fibers(120)
    before = f.current_peak_forces.copy()
    for _ in range(100):
        f.step(np.zeros(120), 1.0)
    assert np.equal(before, f.current_peak_forces).all()
    f = Fibers(120)
    for _ in range(100):
        f.step(np.full(120, 67.0), 1.0)
    assert np.greater(before, f.current_peak_forces).all()
    f = Fibers(120, apply_fatigue=False)
    for _ in range(100):
        f.step(np.full(120, 67.0), 1.0)
    assert np.equal(before, f.current_peak_forces).all()


def test_potvin_muscle():
    from pymuscle_b200 import PotvinFuglevandMuscle as Muscle
    with pytest.raises(TypeError):
        Muscle()
    m = Muscle(120)
    assert m.motor_unit_count == 120
    with pytest.raises(TypeError):
        m.step()
    with pytest.raises(AssertionError):
        m.step(np.ones(3), 1)
    assert m.step(np.zeros(120), 1.0) == pytest.approx(0.0)
    assert Muscle(120).step(np.full(120, 40.0), 1.0) == pytest.approx(1311.86896)
    assert Muscle(120).step(40.0, 1.0) == pytest.approx(1311.86896)
    m = Muscle(120)
    assert m.step(np.full(120, m.max_excitation), 1.0) == pytest.approx(2215.98114)
    assert Muscle(120).step(np.full(120, 107.0), 1.0) == pytest.approx(2215.98114)


def test_standard_muscle():
    from pymuscle_b200 import StandardMuscle as Muscle
    m = Muscle(32.0)
    assert m.motor_unit_count == 120 and m.max_arb_output == pytest.approx(2609.0309)
    m90 = Muscle(90.0)
    assert m90.motor_unit_count == 340 and m90.max_arb_output == pytest.approx(7338.29062)
    with pytest.raises(TypeError):
        m.step()
    with pytest.raises(AssertionError):
        m.step(np.ones(3), 1)
    assert m.step(np.zeros(120), 1.0) == pytest.approx(0.0)
    assert Muscle(32.0).step(np.full(120, 0.5), 1.0) == pytest.approx(0.39096348)
    assert Muscle(32.0).step(0.5, 1.0) == pytest.approx(0.39096348)
    assert Muscle(32.0).step(np.full(120, 1.0), 1.0) == pytest.approx(0.84935028)
    assert Muscle(32.0).step(np.full(120, 41.0), 1.0) == pytest.approx(0.84935028)
    with pytest.raises(NotImplementedError):
        m.get_central_fatigue()


def test_standard_muscle_fatigue():
    from pymuscle_b200 import StandardMuscle as Muscle
    m = Muscle(32.0, apply_peripheral_fatigue=False)
    assert m.get_peripheral_fatigue() == 0.0
    for _ in range(100):
        out = m.step(0.5, 1.0)
    assert out == pytest.approx(0.39096348) and m.get_peripheral_fatigue() == 0.0
    m = Muscle(32.0, apply_peripheral_fatigue=True)
    assert m.get_peripheral_fatigue() == 0.0
    for _ in range(100):
        out = m.step(0.0, 1.0)
    assert out == pytest.approx(0.0) and m.get_peripheral_fatigue() == 0.0
    for _ in range(100):
        out = m.step(0.5, 1.0)
    assert out == pytest.approx(0.29151827)
    assert m.get_peripheral_fatigue() == pytest.approx(0.18028181)


def test_readme_loop_config1():
    """BASELINE config #1: StandardMuscle(), 0.6 excitation, 50 fps, 10,000 steps;
    history of current_forces kept in a list as the README does (README.md:122-130)."""
    from pymuscle_b200 import StandardMuscle
    g = golden("readme_exact")
    muscle = StandardMuscle()
    outputs, history = [], []
    for _ in range(10000):
        outputs.append(muscle.step(0.6, 1 / 50.0))
        if len(outputs) in (1, 100, 10000):
            history.append(muscle.current_forces)
    np.testing.assert_allclose(outputs, g["standard_totals"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(history[-1], g["standard_forces"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(muscle._fibers.current_peak_forces, g["standard_cpf"], rtol=1e-9, atol=1e-10)
    assert history[0] is not history[1] and not np.array_equal(history[0], history[1])
    assert muscle.get_peripheral_fatigue() == pytest.approx(float(g["standard_fatigue"]), rel=1e-9)
    # SURVEY.md Appendix D quotes these for the same run
    assert outputs[0] == pytest.approx(0.5059615150545111, rel=1e-9)
    assert outputs[9999] == pytest.approx(0.1435819929896607, rel=1e-9)
    assert muscle.get_peripheral_fatigue() == pytest.approx(0.5413389785180036, rel=1e-9)


def test_physio_example_loop():
    """examples/minimal-physio-example.py: PotvinFuglevandMuscle(120), 40.0, 10,000 steps."""
    from pymuscle_b200 import PotvinFuglevandMuscle
    g = golden("readme_exact")
    muscle = PotvinFuglevandMuscle(120)
    outputs = [muscle.step(40.0, 0.02) for _ in range(10000)]
    np.testing.assert_allclose(outputs, g["potvin_totals"], rtol=1e-9)
    np.testing.assert_allclose(muscle._fibers.current_peak_forces, g["potvin_cpf"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(muscle._pool._recruitment_durations, g["potvin_dur"], rtol=1e-12)
    assert outputs[0] == pytest.approx(1311.8689626905448, rel=1e-9)
    assert outputs[9999] == pytest.approx(403.39412010089995, rel=1e-9)


def test_generic_composition_with_a_user_model():
    """Muscle(pool, fibers) accepts any Model pair (muscle.py:42-51)."""
    from pymuscle_b200 import Model, Muscle, PotvinFuglevand2017MuscleFibers as Fibers

    class ConstantPool(Model):
        max_excitation = 1.0

        def step(self, inputs, step_size):
            return np.full(self.motor_unit_count, 40.0)

    m = Muscle(ConstantPool(120), Fibers(120))
    assert m.step(0.3, 1.0) == pytest.approx(2586.7530897)
    with pytest.raises(AssertionError):
        Muscle(ConstantPool(10), Fibers(120))


def test_deepcopy_and_pickle_of_muscles():
    """The reference's muscles are plain Python objects: environments deep-copy and pickle them.
    A copy carries the full state and then evolves independently."""
    import copy
    import pickle
    import torch
    import pymuscle_b200 as pm
    m = pm.StandardMuscle()
    for _ in range(30):
        m.step(0.7, 0.05)
    twin = copy.deepcopy(m)
    back = pickle.loads(pickle.dumps(m))
    assert np.array_equal(twin.current_forces, m.current_forces)
    assert twin.get_peripheral_fatigue() == m.get_peripheral_fatigue() == back.get_peripheral_fatigue()
    a, b, c = m.step(0.4, 0.05), twin.step(0.4, 0.05), back.step(0.4, 0.05)
    assert a == b == c
    assert np.array_equal(m.current_forces, twin.current_forces)
    twin.step(1.0, 0.05)                                           # independent from here on
    assert m.step(0.0, 0.05) == back.step(0.0, 0.05)
    assert twin.get_peripheral_fatigue() != m.get_peripheral_fatigue()
    p = pm.PotvinFuglevandMuscle(60)
    p.step(40.0, 0.1)
    q = copy.deepcopy(p)
    assert p.step(40.0, 0.1) == q.step(40.0, 0.1)
    from pymuscle_b200 import MuscleBatch, MuscleSpec
    bt = MuscleBatch(MuscleSpec.potvin(24), 5, "cuda:0", "f32")
    bt.step(torch.full((5,), 30.0, device="cuda:0"), 0.02)
    bt2 = copy.deepcopy(bt)
    assert torch.equal(bt.cpf, bt2.cpf) and bt.cpf.data_ptr() != bt2.cpf.data_ptr()
