"""CPU tests of the trainer's host logic: the REINFORCE update, the flat gradient all-reduce over gloo (world_size 2), the
t-test gate and the checkpoint keys.  (The rollout itself needs the CUDA env: tests/test_gpu_training.py.)"""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dcmrta_b200.policy import AttentionNet
from dcmrta_b200.trainer import agreed_update_count, allreduce_gradients, broadcast_parameters, paired_ttest_improved, reinforce_update


def _batch(seed, n, A=6, T1=9):
    g = torch.Generator().manual_seed(seed)
    tasks = torch.rand(n, T1, 5, generator=g)
    agents = torch.rand(n, A, 6, generator=g)
    mask = torch.rand(n, T1, generator=g) < 0.3
    mask[:, 0] = True
    mask[:, 1] = False
    action = (~mask).float().multinomial(1, generator=g).squeeze(1)
    adv = torch.randn(n, generator=g)
    return tasks, agents, mask, action, adv


def _make(seed=3):
    torch.manual_seed(seed)
    net = AttentionNet(6, 5, 16)
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    sch = torch.optim.lr_scheduler.StepLR(opt, step_size=2000, gamma=0.98)
    return net, opt, sch


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    net, opt, sch = _make(seed=3 + rank)                 # different initial weights on purpose ...
    broadcast_parameters(net)                            # ... rank 0's win
    for it in range(3):
        b = _batch(100 * it + rank, 32)
        reinforce_update(net, opt, sch, *b)
    flat = torch.cat([p.data.flatten() for p in net.parameters()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        out.put([g.numpy() for g in gathered])
    dist.destroy_process_group()


def test_two_rank_update_equals_single_process_on_the_joint_batch():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    g0, g1 = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(g0, g1)                        # the ranks stay in lock-step
    # single process, same weights, the two ranks' mini-batches concatenated: mean over 64 == mean of the two means over 32
    net, opt, sch = _make(seed=3)
    for it in range(3):
        b0, b1 = _batch(100 * it, 32), _batch(100 * it + 1, 32)
        reinforce_update(net, opt, sch, *[torch.cat([x, y]) for x, y in zip(b0, b1)])
    ref = torch.cat([p.data.flatten() for p in net.parameters()]).numpy()
    np.testing.assert_allclose(g0, ref, rtol=2e-4, atol=2e-6)


def _worker_unequal(rank, world, port, out):
    """The shape of ReinforceTrainer.iteration()'s update loop with UNEQUAL decision counts per rank: the ranks
    must issue the same number of gradient all-reduces, then a trailing collective (the all_gather of _eval) must still pair up."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    net, opt, sch = _make(seed=3)
    broadcast_parameters(net)
    n_local = [100, 37][rank]                               # rank-local decision counts: 100 // 16 = 6 updates vs 37 // 16 = 2
    n_up = agreed_update_count(n_local, 16)
    data = _batch(7 + rank, n_local)
    perm = torch.randperm(n_local, generator=torch.Generator().manual_seed(rank))
    for u in range(n_up):
        sel = perm[(u * 16) % n_local:][:16]
        reinforce_update(net, opt, sch, *[x[sel] for x in data])
    none = agreed_update_count([0, 50][rank], 16)           # one rank without decisions: nobody updates
    flat = torch.cat([p.data.flatten() for p in net.parameters()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)                         # would hang or mis-pair if the update counts had differed
    if rank == 0:
        out.put((n_up, none, [g.numpy() for g in gathered]))
    dist.destroy_process_group()


def test_update_count_is_agreed_across_ranks_with_unequal_decision_counts():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_unequal, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    n_up, none, (g0, g1) = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert n_up == 2 and none == 0                          # min(100, 37) // 16
    assert np.array_equal(g0, g1)
    assert agreed_update_count(100, 16) == 6 and agreed_update_count(100, 16, 3) == 3 and agreed_update_count(0, 16, 3) == 0


def test_update_moves_only_live_parameters_and_steps_the_schedule():
    net, opt, sch = _make()
    before = {k: v.clone() for k, v in net.state_dict().items()}
    st = reinforce_update(net, opt, sch, *_batch(0, 64))
    assert all(torch.isfinite(st[k]) for k in ("policy_loss", "entropy", "grad_norm"))
    assert sch.last_epoch == 1
    after = net.state_dict()
    moved = [k for k in before if not torch.equal(before[k], after[k])]
    assert any(k.startswith("pointer.") for k in moved) and any(k.startswith("task_embedding.") for k in moved)
    assert not any("dec_self_attn" in k for k in moved)      # never used by the forward pass (reference attention.py:216-221)
    allreduce_gradients(net)                                  # no process group: a no-op


def test_ttest_gate():
    rng = np.random.default_rng(0)
    base = -40 + rng.normal(0, 3, 256)
    assert paired_ttest_improved(base + 1.0 + rng.normal(0, 0.5, 256), base)[0]
    assert not paired_ttest_improved(base - 1.0, base)[0]                       # worse on average
    assert not paired_ttest_improved(base + rng.normal(0.01, 3, 256), base)[0]  # not significant
