"""GPU tests of the batched rollout and the REINFORCE trainer."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_rollout_buffers_replay_to_the_same_episodes():
    """The experience a rollout records is exactly what the env showed and did: replaying the recorded actions (and leaders)
    on a second env with the same instances reproduces every observation, mask and the episode metrics."""
    from dcmrta_b200 import BatchedTaskEnv
    from dcmrta_b200.policy import AttentionNet
    from dcmrta_b200.rollout import BatchedRollout, clone_instances
    B, A, T = 300, 10, 20
    torch.manual_seed(0)
    net = AttentionNet(6, 5, 32).cuda()
    env = BatchedTaskEnv(B, A, T, auto_reset=False, seed=5)
    env.generate()
    ro = BatchedRollout(env, horizon=4 * (A + T), record=True, check_every=8)
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    ep = ro.run(net, "sample", g)
    assert bool(ep.ended.all()) and ep.length <= ro.horizon
    assert torch.isfinite(ep.reward).all() and (ep.reward < 0).all()
    assert torch.equal(ep.reward, -ep.metrics[:, 2])                       # reward = -makespan (task_env.py:424)
    n_dec = ep.active.sum(0)
    assert torch.equal(n_dec.double(), ep.metrics[:, 7])                   # decisions counted by the kernel == active slots
    # chosen actions were never masked
    chosen_masked = ep.mask.gather(2, ep.action.long().unsqueeze(2)).squeeze(2).bool() & ep.active
    assert not bool(chosen_masked.any())
    # replay
    env2 = BatchedTaskEnv(B, A, T, auto_reset=False, seed=5)
    clone_instances(env, env2)
    env2.reset(leaders=ep.leader[0])
    for t in range(ep.length):
        act = ep.active[t]
        assert torch.equal(env2.agent_obs[act], ep.agent_obs[t][act]), t
        assert torch.equal(env2.task_obs[act], ep.task_obs[t][act]), t
        assert torch.equal(env2.mask_u8[act], ep.mask[t][act]), t
        nxt = ep.leader[t + 1] if t + 1 < ep.length else torch.full_like(ep.leader[0], -1)
        # followers are drawn from the same Philox stream (same seed / gid / decision index), leaders are injected
        env2.step(ep.action[t], next_leaders=torch.where(ep.active[t + 1] if t + 1 < ep.length else torch.zeros_like(act), nxt, torch.full_like(nxt, -1)))
    assert torch.equal(env2.episode_metrics(), ep.metrics)
    # greedy rollouts are deterministic given the instances
    r1 = BatchedRollout(env2, ro.horizon, record=False).run(net, "greedy").reward
    r2 = BatchedRollout(env, ro.horizon, record=False).run(net, "greedy").reward
    assert torch.equal(r1, r2)
    env.close(); env2.close()


def test_trainer_iteration_and_checkpoint_roundtrip(tmp_path):
    from dcmrta_b200.trainer import ReinforceTrainer, TrainerConfig
    cfg = TrainerConfig(agents=10, tasks=20, envs_per_rank=256, batch_size=512, embedding_dim=32, lr=1e-3, eval_instances=64, seed=3)
    tr = ReinforceTrainer(cfg)
    w0 = {k: v.clone() for k, v in tr.net.state_dict().items()}
    out = tr.iteration()
    assert out["episodes"] == 256 and out["updates"] >= 1 and out["decisions"] > 256 * 10
    for k in ("reward", "baseline_reward", "policy_loss", "entropy", "grad_norm", "makespan", "success_rate"):
        assert np.isfinite(out[k]), k
    assert out["reward"] < 0 and 0 <= out["success_rate"] <= 1
    assert any(not torch.equal(w0[k], v) for k, v in tr.net.state_dict().items())
    ev = tr.maybe_update_baseline()
    assert np.isfinite(ev["test_value"]) and 0 <= ev["p"] <= 1
    path = tmp_path / "checkpoint.pth"
    tr.save(path)
    ck = torch.load(path)
    assert set(ck) == {"model", "optimizer", "episode", "lr_decay", "level", "best_perf"}      # driver.py:192-199
    tr2 = ReinforceTrainer(cfg)
    tr2.load(path)
    for (k, a), (_, b) in zip(tr.net.state_dict().items(), tr2.net.state_dict().items()):
        assert torch.equal(a, b), k
    assert tr2.episode == tr.episode == 256
    out2 = tr2.iteration()
    assert np.isfinite(out2["policy_loss"])
