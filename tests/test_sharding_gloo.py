"""CPU test of the multi-process path (world_size 2, gloo): the env partition and the job-total reduction bench.py uses."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dcmrta_b200.sharding import reduce_job_totals, shard_range


def test_shard_range_is_a_partition():
    for total in (0, 1, 7, 65536, 1048576, 1000003):
        for world in (1, 2, 3, 4, 8):
            cover, prev_end = 0, 0
            sizes = []
            for r in range(world):
                first, count = shard_range(total, r, world)
                assert first == prev_end
                prev_end = first + count
                cover += count
                sizes.append(count)
            assert cover == total and max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = shard_range(1000003, rank, world)
    # each rank "steps" its shard 10 times, rank 1 is slower
    steps, seconds = count * 10, 1.0 + rank
    tot_steps, tot_time = reduce_job_totals(steps, seconds)
    ids = torch.zeros(world, 2, dtype=torch.int64)
    ids[rank] = torch.tensor([first, count])
    dist.all_reduce(ids)
    if rank == 0:
        out.put((tot_steps, tot_time, ids.tolist()))
    dist.destroy_process_group()


def test_two_rank_job_totals_over_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    tot_steps, tot_time, ids = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert tot_steps == 1000003 * 10 and tot_time == 2.0
    assert ids[0][0] == 0 and ids[1][0] == ids[0][1] and ids[0][1] + ids[1][1] == 1000003
