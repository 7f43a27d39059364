"""dcmrta_b200/trainer.py -- REINFORCE with a greedy-rollout baseline on device-resident env batches (reference driver.py:61-305,
runner.py, worker.py:87-101), one process per GPU.

What stays exactly as in the reference: loss = -(log pi(a|s) * advantage).mean() (driver.py:163-168), entropy diagnostic (:165),
gradient clipping at L2 norm 10 (:174), Adam(lr=LR) + StepLR(DECAY_STEP, 0.98) stepped per update (:64-65, :175-176), mini-batches
of BATCH_SIZE decisions (:133-138), advantage = episode reward - reward of a greedy episode on the same instance (worker.py:89-94;
played by the CURRENT network as the reference does, TrainerConfig.baseline_net), baseline replaced after a one-sided paired t-test at p < 0.05 on 256 held-out instances (:219-279), and the
checkpoint keys {model, optimizer, episode, lr_decay, level, best_perf} (:192-199) so checkpoints interchange with RL_test.py.

What changes (B200-first): the 8 Ray CPU actors become env shards on the GPUs (sharding.shard_range, no data-path collective);
a rank plays `envs_per_rank` episodes at once through dcmrta_b200.rollout; the only collective is ONE all-reduce per update over
a single flat gradient buffer (~2.5 M fp32 parameters = 10 MB: latency-bound on NVLink 5, so one bucket, no overlap machinery),
issued through torch.distributed (NCCL on GPUs, gloo in the CPU tests).  Every rank draws its own mini-batch, so the effective
batch is world_size x BATCH_SIZE.
"""
from __future__ import annotations

import copy
from dataclasses import dataclass, field

import numpy as np
import torch
import torch.distributed as dist

from .policy import AttentionNet

# reference parameters.py
LR = 1e-5
GAMMA = 1
DECAY_STEP = 2e3
BATCH_SIZE = 1024
EMBEDDING_DIM = 128
AGENT_INPUT_DIM, TASK_INPUT_DIM = 6, 5
MAX_TIME = 100
COALITION_SIZE = 5
EVAL_INSTANCES = 256


def world():
    return (dist.get_rank(), dist.get_world_size()) if dist.is_available() and dist.is_initialized() else (0, 1)


def broadcast_parameters(module: torch.nn.Module, src: int = 0) -> None:
    """every rank starts from rank `src`'s weights (the reference ships global weights to each runner, runner.py:61-62)"""
    if world()[1] == 1:
        return
    flat = torch._utils._flatten_dense_tensors([p.data for p in module.parameters()])
    dist.broadcast(flat, src)
    for p, q in zip(module.parameters(), torch._utils._unflatten_dense_tensors(flat, [p.data for p in module.parameters()])):
        p.data.copy_(q)


def allreduce_gradients(module: torch.nn.Module) -> None:
    """mean of the ranks' gradients, one collective over one flat bucket.  Parameters without a gradient (dec_self_attn, dead in
    the reference as well) are skipped on every rank alike."""
    w = world()[1]
    if w == 1:
        return
    grads = [p.grad for p in module.parameters() if p.grad is not None]
    flat = torch._utils._flatten_dense_tensors(grads)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.div_(w)
    for g, q in zip(grads, torch._utils._unflatten_dense_tensors(flat, grads)):
        g.copy_(q)


def agreed_update_count(n_local: int, batch_size: int, updates_per_iteration: int = 0) -> int:
    """Optimiser steps of this iteration, IDENTICAL on every rank: each step is one gradient all-reduce, and the number of decisions a
    rank collected (n_local) differs from rank to rank, so the ranks agree on min(n_local) first (one scalar all_reduce MIN).  0 when
    some rank has nothing to train on (then nobody updates).  `updates_per_iteration` > 0 overrides the count, not the agreement."""
    n_min = int(n_local)
    if world()[1] > 1:
        t = torch.tensor([n_min], dtype=torch.int64, device="cuda" if dist.get_backend() == "nccl" else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        n_min = int(t.item())
    if n_min == 0:
        return 0
    return int(updates_per_iteration) or max(1, n_min // int(batch_size))


def reinforce_update(net, optimizer, lr_decay, tasks, agents, mask, action, advantage, max_norm: float = 10.0):
    """One optimiser step on one mini-batch (driver.py:162-176).  action [N] int64, advantage [N] fp32.  Returns diagnostics."""
    logp_list = net(tasks, agents, mask)
    logp = torch.gather(logp_list, 1, action.view(-1, 1))
    entropy = (logp_list * logp_list.exp()).nansum(dim=-1).mean()
    policy_loss = (-logp * advantage.view(-1, 1).detach()).mean()
    optimizer.zero_grad(set_to_none=True)
    policy_loss.backward()
    allreduce_gradients(net)
    grad_norm = torch.nn.utils.clip_grad_norm_(net.parameters(), max_norm=max_norm, norm_type=2)
    optimizer.step()
    lr_decay.step()
    return {"policy_loss": policy_loss.detach(), "entropy": entropy.detach(), "grad_norm": grad_norm.detach()}


def paired_ttest_improved(test_value: np.ndarray, baseline_value: np.ndarray, alpha: float = 0.05):
    """driver.py:258-262: better on average AND ttest_rel p < alpha."""
    from scipy.stats import ttest_rel
    if not test_value.mean() > baseline_value.mean():
        return False, 1.0
    _, p = ttest_rel(test_value, baseline_value)
    return bool(p < alpha), float(p)


@dataclass
class TrainerConfig:
    agents: int = 20
    tasks: int = 50
    envs_per_rank: int = 4096
    horizon: int = 0                 # decisions per episode kept in the buffer; 0 = 4 * (agents + tasks)
    batch_size: int = BATCH_SIZE
    lr: float = LR
    decay_step: int = int(DECAY_STEP)
    embedding_dim: int = EMBEDDING_DIM
    max_coalition: int = COALITION_SIZE
    max_time: float = MAX_TIME
    updates_per_iteration: int = 0   # 0 = one pass over the collected decisions
    amp: bool | str = False          # True: rollouts call a bf16 shadow copy of the network (rollout._forward_net); "fused": the inference path of
                                     # policy_fused.py (bf16 GEMMs + the sm_100a kernels of include/dcmrta_policy.h); the update stays fp32
    # Which network plays the greedy episode the advantage is measured against.  "local" is what the reference DOES: baseline_test
    # (worker.py:222) calls self.local_net -- the network that just sampled -- and never the `local_baseline` it was handed, so the
    # advantage is reward - greedy reward of the CURRENT policy and the t-test baseline swap (driver.py:219-279) never reaches the loss.
    # "frozen" is what the reference's variable names suggest: the separately kept baseline network.
    baseline_net: str = "local"
    graph_rollout: bool = True       # replay the decision loop from a CUDA graph (rollout.GraphedRollout); False = eager loop, explicit generator
    # graphed rollouts forward only the envs that are still playing, at these fractions of the batch (GraphedRollout.fractions); (1.0,) = always all.
    # Episode lengths of one batch spread by +-7 % (20A/50T: 120 +- 9 decisions, the longest of 8,192 near 150): with these eight sizes and a poll
    # every 8 decisions 96 % of the forwarded rows are live (79 % without, 92 % with four sizes; simulated on 3,000 oracle episodes).
    rollout_fractions: tuple = (1.0, 0.85, 0.7, 0.5, 0.35, 0.2, 0.1, 0.03)
    seed: int = 0
    eval_instances: int = EVAL_INSTANCES


class ReinforceTrainer:
    def __init__(self, cfg: TrainerConfig, device: int = 0):
        from .batched_env import BatchedTaskEnv
        from .rollout import BatchedRollout, GraphedRollout
        from .sharding import shard_range
        self.cfg = cfg
        self.rank, self.world = world()
        self.device = torch.device("cuda", device)
        torch.manual_seed(cfg.seed)                                  # same initial weights on every rank; broadcast makes it certain
        self.net = AttentionNet(AGENT_INPUT_DIM, TASK_INPUT_DIM, cfg.embedding_dim).to(self.device)
        broadcast_parameters(self.net)
        self.baseline = copy.deepcopy(self.net)
        self.optimizer = torch.optim.Adam(self.net.parameters(), lr=cfg.lr)
        self.lr_decay = torch.optim.lr_scheduler.StepLR(self.optimizer, step_size=cfg.decay_step, gamma=0.98)
        self.episode, self.level, self.best_perf, self.updates = 0, 0, -100.0, 0
        B = cfg.envs_per_rank
        first_gid, _ = shard_range(B * self.world, self.rank, self.world)
        kw = dict(M=cfg.max_coalition, device=device, auto_reset=False, max_time=cfg.max_time)
        self.env = BatchedTaskEnv(B, cfg.agents, cfg.tasks, seed=cfg.seed, first_gid=first_gid, **kw)
        self.base_env = BatchedTaskEnv(B, cfg.agents, cfg.tasks, seed=cfg.seed + 1, first_gid=first_gid, **kw)
        horizon = cfg.horizon or 4 * (cfg.agents + cfg.tasks)
        if cfg.graph_rollout:
            def Rollout(*a, **k):
                return GraphedRollout(*a, fractions=cfg.rollout_fractions, **k)
        else:
            Rollout = BatchedRollout
        self.rollout = Rollout(self.env, horizon, record=True, check_every=8)      # every pass past the longest episode is a wasted forward
        self.base_rollout = Rollout(self.base_env, horizon, record=False, check_every=8)
        self.gen = torch.Generator(device=self.device)
        self.gen.manual_seed(cfg.seed * 1000003 + self.rank)
        if cfg.graph_rollout:                                        # the graphed loop samples from the default CUDA generator
            with torch.cuda.device(self.device):
                torch.cuda.manual_seed(cfg.seed * 1000003 + self.rank)
        # held-out instances for the baseline test (driver.py:119): their own generator stream
        E = max(1, cfg.eval_instances // self.world)
        self.eval_env = BatchedTaskEnv(E, cfg.agents, cfg.tasks, seed=cfg.seed + 7919, first_gid=self.rank * E, **kw)
        self.eval_env.generate(max_duration=5.0)
        self.eval_rollout = Rollout(self.eval_env, horizon, record=False, check_every=8)
        self.baseline_value = None

    # ---- one training iteration: play, score against the baseline, update ------------------------------------------------
    def iteration(self):
        from .rollout import clone_instances
        cfg = self.cfg
        self.env.generate(max_duration=5.0)                          # fresh instances (the reference builds a new TaskEnv per episode, worker.py:32)
        clone_instances(self.env, self.base_env)
        ep = self.rollout.run(self.net, "sample", None if cfg.graph_rollout else self.gen, amp=cfg.amp)
        base = self.base_rollout.run(self.net if cfg.baseline_net == "local" else self.baseline, "greedy", amp=cfg.amp)     # worker.py:89, :200-235
        # every episode trains, as in the reference (worker.py:87-101): one the horizon cut is scored -current_time like a MAX_TIME cut
        valid = torch.ones_like(ep.ended)
        adv_env = (ep.reward - base.reward).float()                                                       # worker.py:93
        use = ep.active & valid.unsqueeze(0)
        idx = use.nonzero(as_tuple=False)                            # [N,2] (t, b)
        N = idx.shape[0]
        perm = torch.randperm(N, device=self.device, generator=self.gen)
        # The number of updates must be the SAME on every rank: each update is one gradient all-reduce, and N (the rank's own decision
        # count) differs from rank to rank.  Agree on min N over the ranks; no rank updates if any rank has nothing to train on.
        n_up = self.agreed_updates(N)
        stats = []
        self.net.train()
        for u in range(n_up):
            sel = idx[perm[(u * cfg.batch_size) % N:][:cfg.batch_size]]
            t, b = sel[:, 0], sel[:, 1]
            stats.append(reinforce_update(self.net, self.optimizer, self.lr_decay, ep.task_obs[t, b], ep.agent_obs[t, b],
                                          ep.mask[t, b].view(torch.bool), ep.action[t, b].long(), adv_env[b]))
            self.updates += 1
        self.episode += self.env.B * self.world
        m = ep.metrics[valid]
        out = {"decisions": int(N), "episodes": int(valid.sum()), "updates": len(stats),
               "reward": float(ep.reward[valid].mean()) if bool(valid.any()) else float("nan"),
               "baseline_reward": float(base.reward[valid].mean()) if bool(valid.any()) else float("nan")}
        for k, name in enumerate(("success_rate", "makespan", "time_cost", "waiting_time", "travel_dist", "efficiency"), start=1):
            out[name] = float(m[:, k].mean()) if m.shape[0] else float("nan")
        for k in ("policy_loss", "entropy", "grad_norm"):
            out[k] = float(torch.stack([s[k] for s in stats]).mean()) if stats else float("nan")
        return out

    def agreed_updates(self, n_local: int) -> int:
        return agreed_update_count(n_local, self.cfg.batch_size, self.cfg.updates_per_iteration)

    # ---- baseline test (driver.py:208-279) ------------------------------------------------------------------------------------
    def _eval(self, net):
        r = self.eval_rollout.run(net, "greedy", amp=self.cfg.amp).reward
        if self.world > 1:
            allr = [torch.empty_like(r) for _ in range(self.world)]
            dist.all_gather(allr, r)
            r = torch.cat(allr)
        return r.cpu().numpy()

    def maybe_update_baseline(self):
        if self.baseline_value is None:
            self.baseline_value = self._eval(self.baseline)
        test_value = self._eval(self.net)
        better, p = paired_ttest_improved(test_value, self.baseline_value)
        if better:
            self.baseline.load_state_dict(self.net.state_dict())
            self.best_perf = float(test_value.mean())
            self.eval_env.generate(max_duration=5.0)                 # new test set (driver.py:273)
            self.baseline_value = None
        return {"test_value": float(test_value.mean()), "baseline_value": float(np.mean(self.baseline_value)) if self.baseline_value is not None else float(test_value.mean()),
                "p": p, "updated": better}

    # ---- checkpoints (driver.py:190-201, :280-287; RL_test.py:28-29 reads ['model']) --------------------------------------------
    def state(self):
        return {"model": self.net.state_dict(), "optimizer": self.optimizer.state_dict(), "episode": self.episode,
                "lr_decay": self.lr_decay.state_dict(), "level": self.level, "best_perf": self.best_perf}

    def save(self, path):
        if self.rank == 0:
            torch.save(self.state(), path)

    def load(self, path_or_state):
        ck = torch.load(path_or_state, map_location=self.device) if not isinstance(path_or_state, dict) else path_or_state
        self.net.load_state_dict(ck["model"])
        self.baseline.load_state_dict(ck["model"])                   # driver.py:79-80
        if "optimizer" in ck:
            self.optimizer.load_state_dict(ck["optimizer"])
        if "lr_decay" in ck:
            self.lr_decay.load_state_dict(ck["lr_decay"])
        self.episode, self.level, self.best_perf = ck.get("episode", 0), ck.get("level", 0), ck.get("best_perf", -100.0)
