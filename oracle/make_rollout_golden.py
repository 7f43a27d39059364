#!/usr/bin/env python
"""oracle/make_rollout_golden.py -- TEST INFRASTRUCTURE (build container only: `python -m oracle.make_rollout_golden`).

Runs the UNMODIFIED reference `Worker.run_episode` (/root/reference/worker.py:41-112, with its `baseline_test` :200-235) on three seeded
10-agent / 20-task instances with fixed reference-`AttentionNet` weights (embedding 16: small fixture, same architecture) and writes
tests/golden/rollout_golden.npz:

  w/<key>                    weights of the network.  (A different `local_baseline` is handed to the Worker, which never uses it:
                             baseline_test (worker.py:222) plays its greedy episode with self.local_net.)
  inst/<k>                   the three instances, [3, ...]
  ep<i>/leader, action, followers, nfol      what np.random.choice / Categorical.sample / env.random_choice drew at every decision
  ep<i>/agents, tasks, mask  episode_buffer slots 0, 1, 3 (worker.py:77-80): the policy inputs of every decision
  ep<i>/buf_action, agent_id, reward, adv    slots 2, 5, 4, 6 (adv after the discount of worker.py:97-101, GAMMA = 1)
  ep<i>/logp                 the reference network's log-probabilities at every decision
  ep<i>/reward, greedy_reward, perf          get_episode_reward, the baseline_test reward, perf_metrics (worker.py:87-108)
  base<i>/leader, action, followers, nfol    the greedy baseline episode of the same instance
  base<i>/margin             smallest gap between the best and the second-best log-probability the baseline saw (greedy replays are
                             only meaningful if it is far above the 2e-5 the two network implementations may differ by)

tests/test_gpu_training.py replays these through dcmrta_b200.rollout.BatchedRollout (choices injected) and compares every buffer.
Nothing of the reference is modified: the draws are recorded through wrappers installed on the instances / on numpy's module attribute.
"""
from __future__ import annotations

import sys
import warnings
from pathlib import Path

import numpy as np
import torch

from . import ref_shim as R

ROOT = Path(__file__).resolve().parent.parent
MAXF = 8


def main():
    warnings.filterwarnings("ignore")
    assert R.available()
    R._stub_matplotlib()
    sys.path.insert(0, str(R.REF_ROOT))
    import worker as W                                       # the reference module, unmodified
    from attention import AttentionNet

    torch.manual_seed(4321)
    net = AttentionNet(6, 5, 16).eval()
    torch.manual_seed(99)
    base = AttentionNet(6, 5, 16).eval()
    out = {}
    for k, v in net.state_dict().items():
        out["w/" + k] = v.numpy().copy()

    insts = []
    real_choice = np.random.choice
    for i, seed in enumerate((3, 14, 27)):
        np.random.seed(1000 + seed)
        torch.manual_seed(2000 + seed)
        w = W.Worker(1, net, base, 0, device="cpu", seed=seed, agents_num=(10, 10), tasks_num=(20, 20))
        insts.append(R.instance_arrays(w.env))
        rec = {"ep": dict(leader=[], followers=[]), "base": dict(leader=[], followers=[], action=[], margin=[])}
        phase = ["ep"]

        def choice(a, *args, **kw):                          # worker.py:54 / :212 leader draws (global numpy RNG)
            r = real_choice(a, *args, **kw)
            rec[phase[0]]["leader"].append(int(r))
            return r

        def follower_recorder(env, key):
            inner = env.random_choice                        # task_env.py:331, drawn from env.rng (the env was seeded)

            def f(a, size=None, replace=True):
                r = inner(a, size, replace) if len(a) else np.array([], dtype=np.int64)
                rec[key]["followers"].append([int(x) for x in np.asarray(r).reshape(-1)])
                return r
            return f

        # every step() of the reference draws followers only when vacancy > 1; record an (empty) entry per decision by wrapping step
        def step_recorder(env, key):
            inner = env.step

            def f(group, leader_id, action, idx=0):
                n0 = len(rec[key]["followers"])
                r = inner(group, leader_id, action, idx)
                if len(rec[key]["followers"]) == n0:
                    rec[key]["followers"].append([])
                if key == "base":
                    rec[key]["action"].append(int(action))
                return r
            return f

        w.env.random_choice = follower_recorder(w.env, "ep")
        w.env.step = step_recorder(w.env, "ep")
        w.baseline_env.random_choice = follower_recorder(w.baseline_env, "base")
        w.baseline_env.step = step_recorder(w.baseline_env, "base")

        # worker.py:200-235 `baseline_test` plays the greedy episode with self.local_net -- the SAME network that sampled, not the
        # `local_baseline` it was handed (which the Worker never uses): the advantage is reward - greedy reward of the current policy.
        inner_net = net.forward

        def net_forward(*a, **kw):                           # margin of every greedy decision
            lp = inner_net(*a, **kw)
            if phase[0] == "base":
                top = torch.topk(lp[0], 2).values
                rec["base"]["margin"].append(float(top[0] - top[1]))
            return lp
        net.forward = net_forward
        inner_bt = w.baseline_test

        def baseline_test():
            phase[0] = "base"
            return inner_bt()
        w.baseline_test = baseline_test
        W.np.random.choice = choice
        try:
            perf = w.run_episode(0)
        finally:
            W.np.random.choice = real_choice
            net.forward = inner_net
        buf = w.experience
        n = len(buf[0])
        assert len(rec["ep"]["leader"]) == n == len(rec["ep"]["followers"]), (len(rec["ep"]["leader"]), n, len(rec["ep"]["followers"]))
        assert len(rec["base"]["leader"]) == len(rec["base"]["followers"]) == len(rec["base"]["action"])
        agents = torch.stack(buf[0]).numpy(); tasks = torch.stack(buf[1]).numpy(); mask = torch.stack(buf[3]).numpy()
        with torch.no_grad():
            logp = net(torch.tensor(tasks), torch.tensor(agents), torch.tensor(mask)).numpy()

        def folarr(lst):
            f = np.full((len(lst), MAXF), -1, np.int8)
            for k, x in enumerate(lst):
                f[k, :len(x)] = x
            return f, np.array([len(x) for x in lst], np.int8)

        fe, ne = folarr(rec["ep"]["followers"])
        fb, nb = folarr(rec["base"]["followers"])
        reward = float(w.env.get_episode_reward(W.MAX_TIME)[0])          # idempotent; the buffer holds its fp32 rounding (worker.py:91)
        assert np.float32(reward) == np.float32(sum(float(x) for x in buf[4]))
        greedy_reward = float(w.baseline_env.get_episode_reward(W.MAX_TIME)[0])
        out.update({
            f"ep{i}/leader": np.array(rec["ep"]["leader"], np.int8), f"ep{i}/followers": fe, f"ep{i}/nfol": ne,
            f"ep{i}/agents": agents.astype(np.float32), f"ep{i}/tasks": tasks.astype(np.float32), f"ep{i}/mask": mask.astype(np.uint8),
            f"ep{i}/action": torch.stack(buf[2]).numpy().reshape(-1).astype(np.int16), f"ep{i}/agent_id": torch.stack(buf[5]).numpy().reshape(-1).astype(np.int8),
            f"ep{i}/buf_reward": torch.stack(buf[4]).numpy().reshape(-1).astype(np.float32), f"ep{i}/adv": torch.stack(buf[6]).numpy().reshape(-1).astype(np.float32),
            f"ep{i}/logp": logp.astype(np.float32), f"ep{i}/reward": np.float64(reward), f"ep{i}/greedy_reward": np.float64(greedy_reward),
            f"ep{i}/perf": np.array([perf[k] for k in ("success_rate", "makespan", "time_cost", "waiting_time", "travel_dist", "efficiency")], np.float64),
            f"base{i}/leader": np.array(rec["base"]["leader"], np.int8), f"base{i}/action": np.array(rec["base"]["action"], np.int16),
            f"base{i}/followers": fb, f"base{i}/nfol": nb, f"base{i}/margin": np.float64(min(rec["base"]["margin"])),
        })
        assert (out[f"ep{i}/agent_id"] == out[f"ep{i}/leader"]).all()
        print(f"instance seed {seed}: {n} sampled decisions, reward {reward:.4f}; baseline {len(rec['base']['leader'])} decisions, reward {greedy_reward:.4f}, "
              f"adv {reward - greedy_reward:.4f}, min greedy margin {min(rec['base']['margin']):.2e}")
        assert min(rec["base"]["margin"]) > 1e-3, "greedy margin too small for a cross-implementation replay: pick another seed"
    for k in ("task_xy", "depot_xy", "req", "dur"):
        out["inst/" + k] = np.stack([x[k] for x in insts])
    np.savez_compressed(ROOT / "tests" / "golden" / "rollout_golden.npz", **out)
    print("saved", ROOT / "tests" / "golden" / "rollout_golden.npz")


if __name__ == "__main__":
    main()
