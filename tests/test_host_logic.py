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


def test_rollout_sub_batch_sizes_keep_the_forwarded_rows_live():
    """rollout.GraphedRollout forwards only the envs that are still playing, at the captured sizes of TrainerConfig.rollout_fractions, re-chosen
    at a poll every 8 decisions.  Host logic only: episode lengths come from the oracle playing the random policy on synthetic 20A/50T
    instances (120 +- 9 decisions); the schedule of sizes is simulated exactly as GraphedRollout.run walks it.  This is how the default
    fractions were chosen: forwarding every env at every decision keeps < 85 % of the forwarded rows live (400 envs; the longest of
    8,192 episodes is longer still), the default > 93 %
    (measured on 8,192 envs on a B200: 77.1 % and 95.5 %, profiles/r14*)."""
    import numpy as np
    from dcmrta_b200.rollout import rows_for, sub_batch_sizes
    from dcmrta_b200.trainer import TrainerConfig
    from oracle.oracle import OracleEnv, synthetic_instance
    assert sub_batch_sizes(300, (1.0, 0.75, 0.5, 0.25, 0.05)) == [15, 75, 150, 225, 300]
    assert sub_batch_sizes(8, (1.0, 0.85, 0.7, 0.5, 0.35, 0.2, 0.1, 0.03)) == [1, 2, 3, 4, 6, 7, 8]
    assert sub_batch_sizes(64, (1.0,)) == [64] and sub_batch_sizes(64, ()) == [64]
    assert rows_for([15, 75, 150, 225, 300], 76) == 150 and rows_for([15, 75, 150], 15) == 15 and rows_for([8], 1) == 8
    B = 400
    lengths = []
    for s in range(B):
        inst = synthetic_instance(20, 50, 5, 5000 + s)
        o = OracleEnv.make(20, inst["task_xy"], inst["depot_xy"], inst["req"], inst["dur"])
        o.seed(11, gid=s, episode=0)
        o.fused_reset()
        n = 0
        while not o.done and n < 400:
            o.fused_step(o.policy_action(1))
            n += 1
        lengths.append(n)
    L = np.array(lengths)
    assert 100 < L.mean() < 140 and L.max() < 280                     # inside the trainer's horizon of 4 (A + T)

    def live_fraction(fractions, every=8):
        sizes, rows, forwarded, t = sub_batch_sizes(B, fractions), B, 0, 0
        while True:
            forwarded += rows * every
            t += every
            n = int((L > t).sum())
            if n == 0:
                return L.sum() / forwarded
            rows = rows_for(sizes, n)
    full, default = live_fraction((1.0,)), live_fraction(TrainerConfig().rollout_fractions)
    assert full < 0.85 and default > 0.93 and default > live_fraction((1.0, 0.75, 0.5, 0.25)) > full + 0.05
