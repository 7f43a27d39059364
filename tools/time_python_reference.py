#!/usr/bin/env python
"""tools/time_python_reference.py [seconds] -- BASELINE.md config 1 in the BUILD container (the reference tree does not travel to the GPU
box): the unmodified reference TaskEnv replaying testSet_20A_50T_CONDET/env_0.pkl under the uniform-random policy, observation + mask
built at every decision (oracle/ref_shim.run_reference_episode = the loop of worker.py:45-87), one core, for `seconds`; then the C
oracle port (oracle/taskenv_oracle.c, bench.py's cpu_baseline) on the same instance and core.  Writes one JSON line (kept under
profiles/): decisions/s of both and their ratio -- the factor between bench.py's `cpu_baseline` (kind "port") and the Python reference."""
import json, os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
from oracle import ref_shim as R
from oracle.oracle import OracleEnv

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 12.0
assert R.available(), "needs /root/reference (build container)"
n, t0, eps = 0, time.perf_counter(), 0
while time.perf_counter() - t0 < secs:
    env = R.load_pickle(0)
    tr, _, _ = R.run_reference_episode(env, "random", seed=eps)
    n += len(tr["leader"]); eps += 1
py_rate = n / (time.perf_counter() - t0)
ia = R.instance_arrays(R.load_pickle(0))
o = OracleEnv.make(**ia)
o.seed(1234, gid=0, episode=0)
o.rollout_bench(1, 20000, seed=1)
t0 = time.perf_counter()
c = o.rollout_bench(1, 400000, seed=2)
port_rate = c / (time.perf_counter() - t0)
print(json.dumps({"what": "BASELINE config 1: reference Python TaskEnv vs the C oracle port, env_0.pkl, random policy, obs + mask every decision, 1 core",
                  "python_reference_decisions_per_s": py_rate, "python_episodes": eps, "python_decisions": n,
                  "c_port_decisions_per_s": port_rate, "port_over_python": port_rate / py_rate,
                  "host": os.uname().nodename, "cores_used": 1, "seconds": secs}))
