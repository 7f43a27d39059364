#!/usr/bin/env python
"""tools/prof_policy.py [B] [which...] -- where a policy forward goes at rollout batch sizes (torch profiler, CUDA time by kernel).
which: fp32 (the module), autocast (bf16 autocast), shadow (a bf16 copy of the module), fused (policy_fused.FusedPolicy); default all."""
import copy, os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dcmrta_b200.policy import AttentionNet
from dcmrta_b200.policy_fused import FusedPolicy
from torch.profiler import profile, ProfilerActivity
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
which = sys.argv[2:] or ["fp32", "autocast", "shadow", "fused"]
A, T = 20, 50
net = AttentionNet(6, 5, 128).cuda().eval()
tasks = torch.rand(B, T + 1, 5, device="cuda"); agents = torch.rand(B, A, 6, device="cuda")
mask = torch.rand(B, T + 1, device="cuda") < 0.5; mask[:, 0] = True; mask[:, 1] = False
shadow = copy.deepcopy(net).to(torch.bfloat16)
fused = FusedPolicy(net)
fns = {"fp32": lambda: net(tasks, agents, mask),
       "autocast": lambda: net(tasks, agents, mask),
       "shadow": lambda: shadow(tasks.to(torch.bfloat16), agents.to(torch.bfloat16), mask).float(),
       "fused": lambda: fused(tasks, agents, mask)}
with torch.no_grad():
    ref = net(tasks, agents, mask)
for name in which:
    fn = fns[name]
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=name == "autocast"):
        for _ in range(3): out = fn()
        err = float((out.float() - ref).abs()[~mask].max())
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); t0 = time.perf_counter(); ev0.record()
        for _ in range(10): fn()
        ev1.record(); torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
        print(f"{name}: {dt*1e3:.2f} ms per forward of {B} envs (GPU {ev0.elapsed_time(ev1) / 10:.2f} ms) -> {B/dt:.3g} env-steps/s; "
              f"max |logp - fp32 module| over legal actions {err:.3g}", flush=True)
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            fn(); torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=70), flush=True)
