#!/usr/bin/env python
"""tools/obs_trace.py [B] -- per-tile phase stamps of one steady-state k_obs_tile (DCM_PASS_TRACE=1): entry, input copies landed,
rows staged, output copies read."""
import os, sys
os.environ["DCM_PASS_TRACE"] = "1"
import ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dcmrta_b200 import BatchedTaskEnv
from dcmrta_b200._lib import lib, check
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
env = BatchedTaskEnv(B, 20, 50, auto_reset=True, seed=1234)
env.generate(); env.reset()
for _ in range(700): env.step(policy="random")
torch.cuda.synchronize()
NT = (B + 31) // 32
buf = np.zeros(NT * 8, np.uint64)
check(lib().dcm_debug_pass_trace(env._h, buf.ctypes.data_as(C.c_void_p), buf.size))
tr = buf.reshape(NT, 8)
slow = (tr[:, 3] >> np.uint64(63)).astype(bool)
tr = (tr & np.uint64((1 << 63) - 1)).astype(np.int64)
t0 = tr[:, 0].min()
t = (tr - t0) / 1e3
pc = lambda x: np.percentile(x, [0, 10, 50, 90, 99, 100]).round(2)
print("B", B, "tiles", NT, "span us", round(t[:, 3].max(), 1), "tiles copied env by env:", int(slow.sum()))
print("entry                 p0/10/50/90/99/100", pc(t[:, 0]))
print("scalars landed - entry                  ", pc(t[:, 1] - t[:, 0]))
print("copies landed - entry                   ", pc(t[:, 4] - t[:, 0]))
print("rows staged - entry                     ", pc(t[:, 2] - t[:, 0]))
print("outputs read - rows staged  (bulk)      ", pc((t[:, 3] - t[:, 2])[~slow]))
if slow.any(): print("outputs written - rows staged (by env)  ", pc((t[:, 3] - t[:, 2])[slow]))
print("block lifetime                          ", pc(t[:, 3] - t[:, 0]))
