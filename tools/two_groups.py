#!/usr/bin/env python
"""tools/two_groups.py [B] [steps] -- G handles of B/G envs each, every handle stepping on its own stream, against one handle
of B envs: the latency-bound k_step of one group overlaps the bandwidth-bound k_obs_tile of another (what an RL loop with
several env groups gets without any change to the library)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dcmrta_b200 import BatchedTaskEnv
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
K = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
for G in (1, 2, 4):
    envs, streams = [], []
    for g in range(G):
        e = BatchedTaskEnv(B // G, 20, 50, auto_reset=True, seed=1234, first_gid=g * (B // G))
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            e.generate(); e.reset()
        envs.append(e); streams.append(st)
    def run(n):
        for _ in range(n):
            for e, st in zip(envs, streams):
                with torch.cuda.stream(st):
                    e.step(policy="random")
    run(300); torch.cuda.synchronize()
    t0 = time.perf_counter(); run(K); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"groups {G} x {B // G} envs: {B * K / dt:.4e} env-steps/s, {dt / K * 1e6:.1f} us per pass over all groups")
    for e in envs: e.close()
