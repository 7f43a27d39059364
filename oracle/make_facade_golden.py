#!/usr/bin/env python
"""oracle/make_facade_golden.py -- TEST INFRASTRUCTURE (build container only: `python -m oracle.make_facade_golden`).

The reference's INDIVIDUAL-agent test loop, worker.py:159-198 `run_test_IS` (RL_test.py METHOD = 'IA'): every decider of a slot acts on
its own through `agent_step`, in id order, with no location groups, no leader draw and no followers.  Recorded on three bundled
instances by running that loop around the REAL reference env with the attention network replaced by the deterministic greedy-nearest
policy (no random draw is left): per decision the agent id, the action, the clock, a digest of (mask, agent obs, task obs) and of the
full live state; final reward, finished flags and perf_metrics.  -> tests/golden/facade_is.npz (tests/test_gpu_facade.py)."""
from __future__ import annotations

import warnings
from pathlib import Path

import numpy as np

from . import canon
from . import ref_shim as R

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def run_is(env, max_time=R.MAX_TIME):
    tr = dict(agent=[], action=[], now=[], dig_obs=[], dig_state=[])
    while not env.finished and env.current_time < max_time:                   # worker.py:163
        decision_agents, current_time = env.next_decision()
        env.current_time = current_time
        env.task_update()
        env.agent_update()
        for agent_id in decision_agents:                                      # worker.py:170
            agent = env.agent_dic[agent_id]
            if not agent["returned"]:
                m = env.get_unfinished_task_mask()
                m = np.insert(m, 0, False) if np.sum(m) == env.tasks_num else np.insert(m, 0, True)
                ag = np.asarray(env.get_current_agent_status(agent), np.float64)
                tk = np.asarray(env.get_current_task_status(agent), np.float64)
                mask = m.astype(np.uint8)
                action = R.greedy_nearest(mask, tk)
                tr["agent"].append(int(agent_id)); tr["action"].append(action); tr["now"].append(float(env.current_time))
                tr["dig_obs"].append(canon.obs_digest(mask, ag.astype(np.float32), tk.astype(np.float32)))
                tr["dig_state"].append(canon.state_digest(R.canonical_state(env)))
                env.agent_step(int(agent_id), action)                         # worker.py:186
                env.task_update()
                env.agent_update()
        env.finished = env.check_finished()
    reward, fin = env.get_episode_reward(max_time)
    return tr, float(reward), np.asarray(fin, bool)


def main():
    warnings.filterwarnings("ignore")
    out = {}
    ids = (0, 5, 9)
    for i in ids:
        env = R.load_pickle(i)
        tr, reward, fin = run_is(env)
        met = R.reference_metrics(env, fin)
        out[f"{i}/agent"] = np.array(tr["agent"], np.int8); out[f"{i}/action"] = np.array(tr["action"], np.int16)
        out[f"{i}/now"] = np.array(tr["now"], np.float64)
        out[f"{i}/dig_obs"] = np.array(tr["dig_obs"], np.uint64); out[f"{i}/dig_state"] = np.array(tr["dig_state"], np.uint64)
        out[f"{i}/finished"] = fin.astype(np.uint8)
        out[f"{i}/metrics"] = np.array([reward, met["success_rate"], met["makespan"], met["time_cost"], met["waiting_time"], met["travel_dist"], met["efficiency"]])
        print(f"env_{i}: {len(tr['agent'])} individual decisions, reward {reward:.4f}, success {met['success_rate']:.2f}")
    out["instances"] = np.array(ids, np.int32)
    np.savez_compressed(OUT / "facade_is.npz", **out)


if __name__ == "__main__":
    main()
