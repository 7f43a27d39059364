#!/usr/bin/env python
"""tools/sanitizer_smoke.py [B] [decisions] -- the hot path at B = 2,048 for compute-sanitizer (memcheck / initcheck / racecheck):
generate, reset, `decisions` passes with the in-kernel random policy and auto-reset (k_step, k_episode_list beside k_obs_tile), a
host-buffer step, a greedy pass, a regenerate handle, the chunked observation kernel and an export / import round trip.
Run as:  compute-sanitizer --tool memcheck python tools/sanitizer_smoke.py   (tools/sanitize.sh keeps the logs under profiles/)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dcmrta_b200 import BatchedTaskEnv

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
N = int(sys.argv[2]) if len(sys.argv) > 2 else 200
env = BatchedTaskEnv(B, 20, 50, M=5, auto_reset=True, seed=99)
env.generate(); env.reset()
ended = 0
for k in range(N):
    env.step(policy="random")
    if k % 20 == 19:
        ended += int(env.done_u8.sum())
out = {"next_leader": np.empty(B, np.int32), "reward": np.empty(B, np.float32), "done": np.empty(B, np.uint8)}
env.step(policy="greedy")
env.step_host(np.ascontiguousarray(env.used_action.cpu().numpy()), out)
raw = env.export_raw(); env.import_raw(raw)
env.build_obs(env.leader.clamp(min=0))                       # granular observation call (k_obs_tile with an explicit leader)
m = env.compute_metrics()
regen = BatchedTaskEnv(515, 10, 20, M=5, auto_reset=True, regenerate=True, seed=3)   # not a multiple of the tile size; episode kernel writes restarted observations
regen.generate(); regen.reset()
for k in range(120):
    regen.step(policy="random")
big = BatchedTaskEnv(257, 30, 100, M=5, auto_reset=True, seed=4)                      # a shape served by the chunked k_obs
big.generate(); big.reset()
for k in range(60):
    big.step(policy="random")
torch.cuda.synchronize()
print(f"sanitizer smoke ok: {B} envs x {N + 2} decisions, {env.total_steps()} env-steps, {env.total_episodes()} episodes accounted, "
      f"{regen.total_episodes()} regenerated restarts, launches {env.launch_count() + regen.launch_count() + big.launch_count()}")
