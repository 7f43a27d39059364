#!/usr/bin/env python
"""tools/pass_trace.py [B] -- timeline of one steady-state k_pass (DCM_PASS_TRACE=1): when step units finish, how long observe
units wait, how long units with an episode restart take."""
import os, sys
os.environ["DCM_PASS_TRACE"] = "1"
import ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dcmrta_b200 import BatchedTaskEnv
from dcmrta_b200._lib import lib, check
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
A, T = 20, 50
env = BatchedTaskEnv(B, A, T, auto_reset=True, seed=1234)
env.generate(); env.reset()
for _ in range(700): env.step(policy="random")
torch.cuda.synchronize()
NT = (B + 31) // 32; CH = 2 + 5
n = NT * (1 + CH)
buf = np.zeros(n * 4, np.uint64)
check(lib().dcm_debug_pass_trace(env._h, buf.ctypes.data_as(C.c_void_p), buf.size))
tr = buf.reshape(n, 4).astype(np.int64)
t0 = tr[:, 0].min()
take, ready, end = (tr[:, 0] - t0) / 1e3, (tr[:, 1] - t0) / 1e3, (tr[:, 2] - t0) / 1e3
ended = (tr[:, 3] >> 32)
st = slice(0, NT); ob = slice(NT, n)
pc = lambda x: np.percentile(x, [0, 10, 50, 90, 99, 100]).round(1)
print("B", B, "units", n, "span us", round(end.max(), 1))
print("step  take  p0/10/50/90/99/100", pc(take[st]))
print("step  end                     ", pc(end[st]))
print("step  duration                ", pc(end[st] - take[st]))
print("obs   take                    ", pc(take[ob]))
print("obs   wait (ready - take)     ", pc(ready[ob] - take[ob]))
print("obs   ready                   ", pc(ready[ob]))
print("obs   work (end - ready)      ", pc(end[ob] - ready[ob]))
print("obs   end                     ", pc(end[ob]))
chunk = (np.arange(n - NT) % CH)
e = ended[ob]
has = np.array([(e[i] > chunk[i]) for i in range(n - NT)])
print("obs units with an episode restart:", int(has.sum()), "work", pc((end[ob] - ready[ob])[has]) if has.any() else "-")
print("obs units without                :", int((~has).sum()), "work", pc((end[ob] - ready[ob])[~has]))
late = np.argsort(end[ob])[-5:]
for i in late: print("  last obs unit", i, "chunk", chunk[i], "ended", e[i], "take %.1f ready %.1f end %.1f" % (take[ob][i], ready[ob][i], end[ob][i]))
