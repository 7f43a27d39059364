#!/usr/bin/env python
"""tools/timeline.py [B] -- kernel start/end times of a few steady-state passes (torch.profiler / CUPTI)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dcmrta_b200 import BatchedTaskEnv
from torch.profiler import profile, ProfilerActivity
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
use_stream = len(sys.argv) > 2
env = BatchedTaskEnv(B, 20, 50, auto_reset=True, seed=1234)
env.generate(); env.reset()
st = torch.cuda.Stream() if use_stream else torch.cuda.current_stream()
with torch.cuda.stream(st):
    for _ in range(600): env.step(policy="random")
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(4): env.step(policy="random")
        torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
for e in ev:
    print("%-28s start %8.1f  end %8.1f  dur %6.1f" % (e.name[:28], e.time_range.start - t0, e.time_range.end - t0, e.time_range.end - e.time_range.start))
