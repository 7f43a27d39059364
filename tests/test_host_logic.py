"""CPU tests of host-side logic that needs no GPU."""
import numpy as np
import pytest


@pytest.mark.reference
def test_instance_generator_matches_reference_rng_protocol():
    """TaskEnv(seed=s) must be the same instance as the reference's (task_env.py:57-71 draw order)."""
    from dcmrta_b200.task_env import generate_instance
    from oracle import ref_shim as R
    Ref = R.ref_taskenv_class()
    for (ar, tr, M, s) in (((10, 20), (20, 50), 5, 0), ((20, 20), (50, 50), 5, 3), (7, 9, 3, 11)):
        ref = Ref(ar, tr, 1, M, seed=s)
        ia = R.instance_arrays(ref)
        A, xy, dep, req, dur, cost = generate_instance(ar, tr, M, 5, np.random.default_rng(s))
        assert A == ia["A"] and np.array_equal(xy, ia["task_xy"]) and np.array_equal(dep, ia["depot_xy"])
        assert np.array_equal(req, ia["req"]) and np.array_equal(dur, ia["dur"])
    np.random.seed(5)
    ref = Ref((10, 20), (20, 50), 1, 5)              # unseeded: global NumPy state (task_env.py:39-48)
    ia = R.instance_arrays(ref)
    np.random.seed(5)
    A, xy, dep, req, dur, cost = generate_instance((10, 20), (20, 50), 5, 5, None)
    assert A == ia["A"] and np.array_equal(xy, ia["task_xy"]) and np.array_equal(req, ia["req"])


def test_algorithmic_bytes_formula():
    """10,633 B per env-step at 20A/50T/M5, 4,493 at 10A/20T, 39,733 at 50A/200T."""
    import bench
    assert bench.algorithmic_bytes(20, 50) == 10633
    assert bench.algorithmic_bytes(10, 20) == 4493
    assert bench.algorithmic_bytes(50, 200) == 39733
