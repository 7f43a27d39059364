#!/usr/bin/env python
"""tools/prof_policy.py [B] -- where a policy forward goes at rollout batch sizes (torch profiler, CUDA time by kernel)."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dcmrta_b200.policy import AttentionNet
from torch.profiler import profile, ProfilerActivity
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
A, T = 20, 50
net = AttentionNet(6, 5, 128).cuda().eval()
tasks = torch.rand(B, T + 1, 5, device="cuda"); agents = torch.rand(B, A, 6, device="cuda")
mask = torch.rand(B, T + 1, device="cuda") < 0.5; mask[:, 0] = True; mask[:, 1] = False
for amp in (False, True):
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
        for _ in range(3): net(tasks, agents, mask)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(10): net(tasks, agents, mask)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
        print(f"amp={amp}: {dt*1e3:.2f} ms per forward of {B} envs -> {B/dt:.3g} env-steps/s")
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            net(tasks, agents, mask); torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=70))
