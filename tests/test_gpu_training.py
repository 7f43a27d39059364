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


def test_rollout_matches_reference_worker():
    """SURVEY 8(f) row 1 pinned to the reference: the UNMODIFIED Worker.run_episode (worker.py:41-112, baseline_test :200-235) was run in
    the build container on three seeded 10A/20T instances with fixed AttentionNet weights (oracle/make_rollout_golden.py).  The same
    weights go into dcmrta_b200.policy.AttentionNet, the reference's draws (leader, sampled action, followers) are injected into
    BatchedRollout.run, and everything the worker produced must come back: the policy inputs of every decision (episode_buffer slots
    0, 1, 3: bit-exact), action and agent id (slots 2, 5), the log-probabilities (<= 2e-5: two implementations of the same network),
    the episode reward, the greedy baseline episode (our network's argmax must pick the recorded actions) and its reward, the
    advantage (slot 6, fp32) and perf_metrics."""
    from pathlib import Path
    from dcmrta_b200 import BatchedTaskEnv
    from dcmrta_b200.policy import AttentionNet
    from dcmrta_b200.rollout import BatchedRollout
    z = np.load(Path(__file__).resolve().parent / "golden" / "rollout_golden.npz")
    net = AttentionNet(6, 5, 16)
    net.load_state_dict({k[2:]: torch.tensor(z[k]) for k in z.files if k.startswith("w/")}, strict=True)
    net = net.cuda()
    B, A, T = 3, 10, 20
    dev = torch.device("cuda", 0)

    def replay_of(prefix, with_action):
        n = [len(z[f"{prefix}{i}/leader"]) for i in range(B)]
        L = max(n)
        leader = torch.full((L + 1, B), -1, dtype=torch.int32)
        fol = torch.full((L, B, 8), -1, dtype=torch.int32)
        act = torch.zeros((L, B), dtype=torch.int32)
        for i in range(B):
            leader[:n[i], i] = torch.tensor(z[f"{prefix}{i}/leader"].astype(np.int32))
            fol[:n[i], i] = torch.tensor(z[f"{prefix}{i}/followers"].astype(np.int32))
            act[:n[i], i] = torch.tensor(z[f"{prefix}{i}/action"].astype(np.int32))
        r = {"leader": leader.to(dev), "followers": fol.to(dev)}
        if with_action:
            r["action"] = act.to(dev)
        return r, n, L, act

    def make_env():
        e = BatchedTaskEnv(B, A, T, M=5, auto_reset=False)
        e.load_instances(z["inst/task_xy"], z["inst/depot_xy"], z["inst/req"], z["inst/dur"])
        return e

    # ---- the sampled episode (worker.py:45-85)
    rp, n, L, _ = replay_of("ep", True)
    env = make_env()
    ep = BatchedRollout(env, horizon=L, record=True, check_every=1000).run(net, "sample", replay=rp, keep_logp=True)
    assert ep.length == L and bool(ep.ended.all())
    for i in range(B):
        k = n[i]
        assert bool(ep.active[:k, i].all()) and not bool(ep.active[k:, i].any())
        assert np.array_equal(ep.agent_obs[:k, i].cpu().numpy(), z[f"ep{i}/agents"]), i          # slot 0
        assert np.array_equal(ep.task_obs[:k, i].cpu().numpy(), z[f"ep{i}/tasks"]), i            # slot 1
        assert np.array_equal(ep.mask[:k, i].cpu().numpy(), z[f"ep{i}/mask"]), i                 # slot 3
        assert np.array_equal(ep.action[:k, i].cpu().numpy(), z[f"ep{i}/action"]), i             # slot 2
        assert np.array_equal(ep.leader[:k, i].cpu().numpy(), z[f"ep{i}/agent_id"]), i           # slot 5
        ref_lp, lp = z[f"ep{i}/logp"], ep.logp[:k, i].cpu().numpy()
        free = z[f"ep{i}/mask"] == 0
        np.testing.assert_allclose(lp[free], ref_lp[free], rtol=0, atol=2e-5)
        assert float(ep.reward[i]) == float(z[f"ep{i}/reward"])                                 # worker.py:87, f64 bit-exact
        m = ep.metrics[i].cpu().numpy()
        perf = z[f"ep{i}/perf"]                                                                  # worker.py:103-108
        for c, name in enumerate(("success_rate", "makespan", "time_cost", "waiting_time", "travel_dist", "efficiency")):
            if name == "waiting_time":
                assert m[1 + c] == pytest.approx(perf[c], rel=1e-12), (i, name)
            else:
                assert m[1 + c] == perf[c], (i, name)
    # ---- the greedy baseline episode of the same instances, played by the SAME network (worker.py:200-235)
    rpb, nb, Lb, act_b = replay_of("base", False)
    env_b = make_env()
    base = BatchedRollout(env_b, horizon=Lb, record=True, check_every=1000).run(net, "greedy", replay=rpb)
    for i in range(B):
        assert np.array_equal(base.action[:nb[i], i].cpu().numpy(), act_b[:nb[i], i].numpy()), i  # argmax picks what the reference's argmax picked
        assert float(base.reward[i]) == float(z[f"ep{i}/greedy_reward"])
    # ---- advantage (worker.py:93-101, GAMMA = 1: every decision of the episode carries reward - greedy reward, in fp32) and slot 4
    adv = (ep.reward - base.reward).float().cpu().numpy()
    for i in range(B):
        assert (z[f"ep{i}/adv"] == adv[i]).all(), i
        assert z[f"ep{i}/buf_reward"][-1] == np.float32(float(ep.reward[i])) and not z[f"ep{i}/buf_reward"][:-1].any()
    env.close(); env_b.close()


def test_second_rollout_trains_on_its_first_decision():
    """reset() must clear the done flags of the previous episode, otherwise active[0] of every later rollout is False
    and the first decision of every episode drops out of training."""
    from dcmrta_b200 import BatchedTaskEnv
    from dcmrta_b200.policy import AttentionNet
    from dcmrta_b200.rollout import BatchedRollout
    torch.manual_seed(0)
    net = AttentionNet(6, 5, 16).cuda()
    env = BatchedTaskEnv(64, 6, 10, auto_reset=False, seed=2)
    env.generate()
    ro = BatchedRollout(env, horizon=4 * 16, record=True, check_every=4)
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    first = ro.run(net, "sample", g)
    assert bool(first.active[0].all()) and bool(first.ended.all())
    second = ro.run(net, "sample", g)
    assert bool(second.active[0].all())
    assert torch.equal(second.active.sum(0).double(), second.metrics[:, 7])
    env.close()


def test_horizon_cut_episodes_are_scored_not_dropped():
    """an episode the buffer horizon cuts is scored -current_time (the reference's MAX_TIME cut, worker.py:45/:87), not NaN."""
    from dcmrta_b200 import BatchedTaskEnv
    from dcmrta_b200.policy import AttentionNet
    from dcmrta_b200.rollout import BatchedRollout
    torch.manual_seed(0)
    net = AttentionNet(6, 5, 16).cuda()
    env = BatchedTaskEnv(32, 10, 20, auto_reset=False, seed=4)
    env.generate()
    ep = BatchedRollout(env, horizon=10, record=True).run(net, "sample", torch.Generator(device="cuda").manual_seed(1))
    assert not bool(ep.ended.any()) and torch.isfinite(ep.reward).all()
    assert torch.equal(ep.reward, -env.get_clock())
    env.close()


def _nccl_worker(rank, world, port, q):
    import os
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from dcmrta_b200 import trainer as T
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    calls = []
    real = dist.all_reduce

    def counting(t, *a, **k):
        calls.append(int(t.numel()))
        return real(t, *a, **k)
    dist.all_reduce = counting
    cfg = T.TrainerConfig(agents=6, tasks=10, envs_per_rank=128 + 64 * rank, batch_size=256, embedding_dim=16, lr=1e-3, eval_instances=16, seed=5)
    tr = T.ReinforceTrainer(cfg, device=rank)
    out = tr.iteration()                                   # unequal decision counts per rank: the update count is agreed first
    ev = tr.maybe_update_baseline()                        # the all_gather after the updates must still pair up
    flat = torch.cat([p.data.flatten() for p in tr.net.parameters()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        q.put(dict(identical=all(torch.equal(gathered[0], g) for g in gathered), updates=out["updates"], decisions=out["decisions"],
                   grad_allreduces=sum(1 for c in calls if c > 1000), p=ev["p"]))
    dist.destroy_process_group()


def test_trainer_iteration_two_ranks_nccl():
    """BASELINE configs[4] on hardware: one ReinforceTrainer.iteration() per rank on two GPUs over NCCL, with a different number of envs
    (hence of decisions) per rank.  The ranks must issue the same number of gradient all-reduces and end with bit-identical parameters."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res["identical"] and res["updates"] >= 1 and res["grad_allreduces"] == res["updates"]


def test_graphed_rollout_records_what_the_env_did():
    """rollout.GraphedRollout (the decision loop replayed from a CUDA graph, dcm_step captured with its stream fork / join): what it
    records must replay, decision by decision, on a second env through the eager path."""
    from dcmrta_b200 import BatchedTaskEnv
    from dcmrta_b200.policy import AttentionNet
    from dcmrta_b200.rollout import GraphedRollout, clone_instances
    B, A, T = 300, 10, 20
    torch.manual_seed(0)
    net = AttentionNet(6, 5, 32).cuda()
    env = BatchedTaskEnv(B, A, T, auto_reset=False, seed=5)
    env.generate()
    ro = GraphedRollout(env, horizon=4 * (A + T), record=True, check_every=8, unroll=4)
    torch.cuda.manual_seed(7)
    ep = ro.run(net, "sample")                                 # (the first call also builds the graph: two eager warm-up decisions draw as well)
    assert bool(ep.ended.all())
    n_dec = ep.active.sum(0)
    assert torch.equal(n_dec.double(), ep.metrics[:, 7])
    assert not bool((ep.mask.gather(2, ep.action.long().unsqueeze(2)).squeeze(2).bool() & ep.active).any())
    env2 = BatchedTaskEnv(B, A, T, auto_reset=False, seed=5)
    clone_instances(env, env2)
    env2.reset(leaders=ep.leader[0])
    for t in range(ep.length):
        act = ep.active[t]
        assert torch.equal(env2.agent_obs[act], ep.agent_obs[t][act]), t
        assert torch.equal(env2.task_obs[act], ep.task_obs[t][act]), t
        assert torch.equal(env2.mask_u8[act], ep.mask[t][act]), t
        more = t + 1 < ep.length
        nxt = torch.where(ep.active[t + 1], ep.leader[t + 1], torch.full_like(ep.leader[0], -1)) if more else torch.full_like(ep.leader[0], -1)
        env2.step(ep.action[t], next_leaders=nxt)
    assert torch.equal(env2.episode_metrics(), ep.metrics)
    # the same graph again (the env's Philox streams have moved on to the next episode index, so the episode differs): still a whole,
    # self-consistent record
    again = ro.run(net, "sample")
    assert bool(again.ended.all()) and bool(again.active[0].all())
    assert torch.equal(again.active.sum(0).double(), again.metrics[:, 7])
    assert not bool((again.mask.gather(2, again.action.long().unsqueeze(2)).squeeze(2).bool() & again.active).any())
    env.close(); env2.close()


def test_graphed_rollout_forwards_only_live_envs():
    """GraphedRollout(fractions=...): late in a rollout the policy is called on a gathered list of the envs that are still playing.  The
    record must still replay decision by decision, every chosen action must be legal for ITS env (a misrouted action would hit another
    env's mask), fewer rows than B x length must have been forwarded, and a greedy rollout must choose what the network says for the
    recorded observation of each env, also where the list is in use."""
    from dcmrta_b200 import BatchedTaskEnv
    from dcmrta_b200.policy import AttentionNet
    from dcmrta_b200.rollout import GraphedRollout, clone_instances
    B, A, T = 300, 10, 20
    torch.manual_seed(0)
    net = AttentionNet(6, 5, 32).cuda().eval()
    env = BatchedTaskEnv(B, A, T, auto_reset=False, seed=5)
    env.generate()
    ro = GraphedRollout(env, horizon=4 * (A + T), record=True, check_every=4, unroll=4, fractions=(1.0, 0.75, 0.5, 0.25, 0.05))
    assert ro.sizes == [15, 75, 150, 225, 300]
    torch.cuda.manual_seed(7)
    ep = ro.run(net, "sample")
    assert bool(ep.ended.all())
    assert ep.forwarded < B * ep.length and ep.forwarded >= int(ep.active.sum())
    assert torch.equal(ep.active.sum(0).double(), ep.metrics[:, 7])
    assert not bool((ep.mask.gather(2, ep.action.long().unsqueeze(2)).squeeze(2).bool() & ep.active).any())
    env2 = BatchedTaskEnv(B, A, T, auto_reset=False, seed=5)
    clone_instances(env, env2)
    env2.reset(leaders=ep.leader[0])
    for t in range(ep.length):
        act = ep.active[t]
        assert torch.equal(env2.agent_obs[act], ep.agent_obs[t][act]), t
        assert torch.equal(env2.task_obs[act], ep.task_obs[t][act]), t
        assert torch.equal(env2.mask_u8[act], ep.mask[t][act]), t
        more = t + 1 < ep.length
        nxt = torch.where(ep.active[t + 1], ep.leader[t + 1], torch.full_like(ep.leader[0], -1)) if more else torch.full_like(ep.leader[0], -1)
        env2.step(ep.action[t], next_leaders=nxt)
    assert torch.equal(env2.episode_metrics(), ep.metrics)
    g = ro.run(net, "greedy")
    assert bool(g.ended.all()) and g.forwarded < B * g.length
    agree, total = 0, 0
    with torch.no_grad():
        for t in range(g.length):
            act = g.active[t]
            if 0 < int(act.sum()) <= 150:                                  # decisions taken while a gathered list was in use
                want = net(g.task_obs[t][act], g.agent_obs[t][act], g.mask[t][act].view(torch.bool)).argmax(1).int()
                agree += int((want == g.action[t][act]).sum()); total += int(act.sum())
    assert total > 0 and agree >= 0.995 * total, (agree, total)
    env.close(); env2.close()


def test_bf16_shadow_rollout_follows_the_fp32_weights():
    """amp=True: the decision loop calls a bf16 shadow copy of the network; the copy is refreshed from the fp32 weights at the start of every
    run, so a rollout after an update plays the UPDATED policy (also through the CUDA graph, whose nodes keep the shadow's addresses)."""
    from dcmrta_b200 import BatchedTaskEnv
    from dcmrta_b200.policy import AttentionNet
    from dcmrta_b200.rollout import GraphedRollout
    torch.manual_seed(0)
    net = AttentionNet(6, 5, 32).cuda()
    env = BatchedTaskEnv(128, 10, 20, auto_reset=False, seed=9)
    env.generate()
    ro = GraphedRollout(env, horizon=120, record=True, check_every=8, unroll=4)
    import copy

    def bf16_argmax(n, ep):                                      # what a bf16 copy of network `n` picks at the first decision of episode `ep`
        nb = copy.deepcopy(n).to(torch.bfloat16).eval()
        with torch.no_grad():
            lp = nb(ep.task_obs[0].to(torch.bfloat16), ep.agent_obs[0].to(torch.bfloat16), ep.mask[0].view(torch.bool)).float()
        return lp.argmax(1).int()
    a = ro.run(net, "greedy", amp=True)
    assert bool(a.ended.all()) and not bool((a.mask.gather(2, a.action.long().unsqueeze(2)).squeeze(2).bool() & a.active).any())
    assert torch.equal(bf16_argmax(net, a), a.action[0])          # the loop played the bf16 copy of the CURRENT weights
    old = copy.deepcopy(net)
    with torch.no_grad():
        for p in net.parameters():
            p.add_(torch.randn_like(p) * 0.5)                    # "an update"
    b = ro.run(net, "greedy", amp=True)                          # same graph, same shadow object, refreshed weights
    assert torch.equal(bf16_argmax(net, b), b.action[0])          # the shadow follows the new weights ...
    assert (bf16_argmax(old, b) == b.action[0]).float().mean() < 0.9      # ... not the old ones
    env.close()
