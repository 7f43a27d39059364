"""dcmrta_b200/sharding.py -- env shards across the GPUs of one box.

Envs are independent: the data path has NO collective.  Rank r of W owns the contiguous block of global env ids
[first_gid, first_gid + count); Philox streams are keyed by the GLOBAL id, so the trajectory of env k does not depend on
W or on the GPU it lives on (tests/test_gpu_parity.py::test_shard_invariance).  The only communication is the scalar
reduction of counters / timings for reporting (and, in the trainer, the gradient all-reduce)."""
from __future__ import annotations

import os


def shard_range(total_envs: int, rank: int, world: int):
    """Contiguous, balanced partition: the first (total % world) ranks get one extra env.  -> (first_gid, count)"""
    if not (0 <= rank < world) or total_envs < 0:
        raise ValueError("bad rank / world / total")
    base, extra = divmod(total_envs, world)
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count


def dist_env():
    """(rank, local_rank, world) from the torchrun environment (defaults for a single process)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def reduce_job_totals(steps: float, seconds: float, group=None):
    """Whole-job totals: env-steps summed over ranks, time = max over ranks (the slowest rank bounds the job)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(steps), float(seconds)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    s = torch.tensor([float(steps)], dtype=torch.float64, device=dev)
    t = torch.tensor([float(seconds)], dtype=torch.float64, device=dev)
    dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(s.item()), float(t.item())
