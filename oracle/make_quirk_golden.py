"""oracle/make_quirk_golden.py -- small fixtures that hit the reference's quirks one by one,
recorded by executing the REAL reference (build container only: `python -m oracle.make_quirk_golden`).  TEST INFRASTRUCTURE ONLY.

Writes tests/golden/quirks.npz with the schema of traces_sweep.npz (per-decision leader / action / followers / clock / reward and
digests of the observation and of the full live state, final metrics and digest) plus, per fixture, the instance and a CENSUS of
which quirks the reference actually exercised while it ran (counted on the reference's own dicts by the hooks below):

  hand-built
    q10_groups      four agents finish four single-agent tasks placed symmetrically around the depot at the same instant, so one
                    slot has FOUR location groups; np.unique(axis=0) orders them by (x, then y), not by agent id (task_env.py:291-298)
    q10_pairs       two coalitions of two finish at the same instant at two locations: two groups of two (leader + follower drawn
                    inside each group)
    q11_depot       the whole group follows a leader that picks the depot while tasks are open (task_env.py:327-336); the clock then
                    jumps to the latest arrival when nobody can decide (Q13, :286 / :369)
    q2_q4_wait      a coalition that can never fill waits max_waiting_time and gives up: two members with the same arrival expire in
                    one call, the second is skipped (Q2) and caught by the next call; a complete coalition spread over more than
                    max_waiting_time sheds its early members (Q4); status stays stale after removals (Q3)
  searched        small random instances (2-5 agents, 4-9 tasks) under the random policy, chosen greedily over seeds until every
                  countable quirk is covered at least twice: ghost members (Q1), re-visits (Q8), sticky `assigned` (Q5), ...
"""
from __future__ import annotations

import sys
import warnings
from pathlib import Path

import numpy as np

from . import canon
from . import ref_shim as R
from .make_golden import MAXF, pack

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"
CENSUS_KEYS = ("q1_ghost_member_decides", "q2_skipped_after_removal", "q3_stale_status", "q4_spread_removal", "q5_sticky_assigned",
               "q8_revisit", "q10_multi_group_slot", "q10_max_groups", "q11_group_follows_to_depot", "q13_clock_jump", "decisions")


class Census:
    """Counts quirk occurrences on the reference's own state, through wrappers around its methods (the reference code is not modified)."""

    def __init__(self, env):
        self.env, self.c = env, dict.fromkeys(CENSUS_KEYS, 0)
        self._task_update, self._agent_step, self._next_decision, self._agent_update = env.task_update, env.agent_step, env.next_decision, env.agent_update
        env.task_update, env.agent_step, env.next_decision, env.agent_update = self.task_update, self.agent_step, self.next_decision, self.agent_update

    def agent_update(self):
        env = self.env
        out = self._agent_update()
        for i, a in env.agent_dic.items():                                # Q5: `assigned` left over from the PREVIOUS task while the new one has not started
            if a["route"] and a["route"][-1] >= 0 and a["assigned"]:
                t = env.task_dic[a["route"][-1]]
                if t["feasible_assignment"] and i in t["members"] and env.current_time < t["time_start"]:
                    self.c["q5_sticky_assigned"] += 1
        return out

    def task_update(self):
        env = self.env
        before = {j: (list(t["members"]), bool(t["feasible_assignment"])) for j, t in env.task_dic.items()}
        out = self._task_update()
        W, now = env.max_waiting_time, env.current_time
        for j, t in env.task_dic.items():
            mem0, feas0 = before[j]
            if feas0:
                continue
            removed = len(mem0) - len(t["members"])
            req = int(np.asarray(t["requirements"]).reshape(-1)[0])
            if removed and req - len(mem0) <= 0:
                self.c["q4_spread_removal"] += 1
            if removed and req - len(mem0) > 0 and any(now - env.get_arrival_time(m, j) >= W for m in t["members"]):
                self.c["q2_skipped_after_removal"] += 1                   # somebody who should have gone is still there
            if not t["feasible_assignment"] and int(np.asarray(t["status"]).reshape(-1)[0]) != req - len(t["members"]):
                self.c["q3_stale_status"] += 1
        return out

    def agent_step(self, agent_id, task_id):
        env = self.env
        if task_id != 0 and agent_id in env.task_dic[task_id - 1]["members"]:
            self.c["q8_revisit"] += 1
        return self._agent_step(agent_id, task_id)

    def next_decision(self):
        ids, t = self._next_decision()
        if len(ids) == 0:
            self.c["q13_clock_jump"] += 1
        return ids, t

    def on_slot(self, env, groups):
        if len(groups) > 1:
            self.c["q10_multi_group_slot"] += 1
        self.c["q10_max_groups"] = max(self.c["q10_max_groups"], len(groups))

    def on_decision(self, env, leader, mask, group_size_hint=None):
        a = env.agent_dic[leader]
        self.c["decisions"] += 1
        if a["route"] and a["route"][-1] >= 0:
            t = env.task_dic[a["route"][-1]]
            if not t["feasible_assignment"] and leader in t["members"]:
                self.c["q1_ghost_member_decides"] += 1

    def vector(self):
        return np.array([self.c[k] for k in CENSUS_KEYS], np.int64)


def build_env(A, task_xy, depot_xy, req, dur, M):
    """A reference TaskEnv on a hand-built instance: constructed with the right sizes, then reset(test_env=...) exactly as RL_test.py:36-43."""
    TaskEnv = R.ref_taskenv_class()
    T = len(task_xy)
    env = TaskEnv((A, A), (T, T), 1, M, seed=0)
    for j in range(T):
        env.task_dic[j]["location"] = np.asarray(task_xy[j], np.float64)
        env.task_dic[j]["requirements"] = np.array([int(req[j])])
        env.task_dic[j]["time"] = float(dur[j])
    env.depot["location"] = np.asarray(depot_xy, np.float64)
    for i in range(A):
        env.agent_dic[i]["depot"] = env.depot["location"]
    env.max_waiting_time = 10
    env.reactive_planning = False
    env.reset((env.task_dic, env.agent_dic, env.depot))
    env.clear_decisions()
    return env


def record(env, policy, seed, leader_fn=None):
    cen = Census(env)
    digs_o, digs_s = [], []

    def on_dec(env_, leader, mask, ag, tk):
        st = R.canonical_state(env_)
        digs_o.append(canon.obs_digest(mask, ag.astype(np.float32), tk.astype(np.float32)))
        digs_s.append(canon.state_digest(st))
        cen.on_decision(env_, leader, mask)

    depot_followers = [0]
    step0 = env.step

    def step(group, leader_id, action, idx=0):
        if action == 0 and len(group) > 1:
            depot_followers[0] += 1
        return step0(group, leader_id, action, idx)

    env.step = step
    trace, reward, fin = R.run_reference_episode(env, policy, seed, on_decision=on_dec, leader_fn=leader_fn, on_slot=cen.on_slot)
    cen.c["q11_group_follows_to_depot"] = depot_followers[0]
    met = R.reference_metrics(env, fin)
    n = len(trace["leader"])
    fol = np.full((n, MAXF), -1, np.int8)
    for k, f in enumerate(trace["followers"]):
        fol[k, :len(f)] = f
    ep = dict(leader=np.array(trace["leader"], np.int8), action=np.array(trace["action"], np.int16),
              nfol=np.array([len(f) for f in trace["followers"]], np.int8), followers=fol,
              now=np.array(trace["now"], np.float64), reward=np.array(trace["reward"], np.float64),
              dig_obs=np.array(digs_o, np.uint64), dig_state=np.array(digs_s, np.uint64),
              metrics=np.array([reward, met["success_rate"], met["makespan"], met["time_cost"], met["waiting_time"],
                                met["travel_dist"], met["efficiency"], float(n)], np.float64),
              finished=np.asarray(fin, np.uint8), final_digest=np.array([canon.state_digest(R.canonical_state(env))], np.uint64))
    return ep, cen


def scripted(actions):
    """policy: the k-th decision takes actions[k] when that entry is unmasked, else the first unmasked entry."""
    def pol(env, leader, mask, k):
        a = actions[k] if k < len(actions) else -1
        if 0 <= a < len(mask) and mask[a] == 0:
            return a
        return int(np.flatnonzero(mask == 0)[0])
    return pol


def hand_built():
    fx = []
    lowest = lambda group, k: min(group)
    # ---- Q10: four single-agent tasks at distance 0.25 from the depot in the four axis directions (all coordinates dyadic, so the four
    #      distances, arrivals and finish times are bit-identical); a second ring, asymmetric, for what follows
    ring1 = [(0.75, 0.5), (0.25, 0.5), (0.5, 0.75), (0.5, 0.25)]
    ring2 = [(0.875, 0.625), (0.125, 0.375), (0.625, 0.9375), (0.375, 0.0625)]
    inst = dict(A=4, task_xy=np.array(ring1 + ring2), depot_xy=np.array([0.5, 0.5]), req=np.ones(8, np.int32), dur=np.full(8, 5.0))
    fx.append(("q10_groups", inst, 3, scripted([1, 2, 3, 4, 5, 6, 7, 8]), lowest))
    # ---- Q10 with groups of two: two coalitions of two at mirrored locations, finishing together
    xy = [(0.75, 0.5), (0.25, 0.5), (0.5, 0.875), (0.5, 0.125), (0.9375, 0.75), (0.0625, 0.25)]
    inst = dict(A=4, task_xy=np.array(xy), depot_xy=np.array([0.5, 0.5]), req=np.array([2, 2, 2, 2, 1, 1], np.int32), dur=np.full(6, 5.0))
    fx.append(("q10_pairs", inst, 3, scripted([1, 2, 3, 4, 5, 6]), lowest))
    # ---- Q11 / Q13: agent 0 leaves for a task, then the leader of the remaining group of two picks the depot while tasks are open (legal
    #      for step(); the worker's mask would forbid it) and the other one follows; agent 0 serves the three tasks alone
    inst = dict(A=3, task_xy=np.array([(0.8125, 0.3125), (0.1875, 0.6875), (0.4375, 0.9375)]), depot_xy=np.array([0.5, 0.5]),
                req=np.array([1, 1, 1], np.int32), dur=np.full(3, 5.0))

    def depot_second(env, leader, mask, k):
        return 0 if k == 1 else int(np.flatnonzero(mask == 0)[0])
    fx.append(("q11_depot", inst, 3, depot_second, lowest))
    # ---- the same with the WHOLE team following to the depot at the first decision: nobody can ever decide again and nothing is finished.
    #      The reference loop (worker.py:45) would spin forever; the port stops after two empty slots (DCM_ENV_STUCK)
    def depot_first(env, leader, mask, k):
        return 0
    fx.append(("q11_depot_stuck", inst, 3, depot_first, lowest))
    # ---- Q2 / Q3 / Q4: two agents, tasks that need three: the pair waits out max_waiting_time together (same arrival: Q2 skips the second)
    inst = dict(A=2, task_xy=np.array([(0.6875, 0.8125), (0.3125, 0.1875), (0.9375, 0.0625), (0.0625, 0.5625)]), depot_xy=np.array([0.40625, 0.53125]),
                req=np.array([3, 3, 2, 1], np.int32), dur=np.full(4, 5.0))
    fx.append(("q2_q4_wait", inst, 3, scripted([1, 2, 3, 2, 1, 4, 3]), lowest))
    return fx


def main():
    warnings.filterwarnings("ignore")
    assert R.available(), "reference tree missing"
    eps, names, extra = [], [], {}
    total = np.zeros(len(CENSUS_KEYS), np.int64)

    def keep(name, inst, ep, cen):
        eps.append(ep); names.append(name)
        for k in ("task_xy", "depot_xy", "req", "dur"):
            extra[f"inst/{name}/{k}"] = np.asarray(inst[k])
        extra[f"inst/{name}/A"] = np.int32(inst["A"])
        extra[f"inst/{name}/M"] = np.int32(inst["M"])
        extra[f"census/{name}"] = cen.vector()
        extra[f"finished/{name}"] = ep["finished"]
        print(f"{name:28s} A={inst['A']} T={len(inst['req'])} decisions={len(ep['leader'])}", {k: v for k, v in cen.c.items() if v and k != "decisions"})

    for name, inst, M, pol, lf in hand_built():
        inst = dict(inst, M=M)
        env = build_env(inst["A"], inst["task_xy"], inst["depot_xy"], inst["req"], inst["dur"], M)
        ep, cen = record(env, pol, seed=11, leader_fn=lf)
        keep(name, inst, ep, cen)
        total += cen.vector()
    hb = {n: extra[f"census/{n}"] for n in names}
    ix = CENSUS_KEYS.index
    assert hb["q10_groups"][ix("q10_max_groups")] == 4, "q10_groups did not produce four location groups"
    assert hb["q10_pairs"][ix("q10_multi_group_slot")] >= 1
    assert hb["q11_depot"][ix("q11_group_follows_to_depot")] >= 1 and hb["q11_depot"][ix("q13_clock_jump")] >= 1
    assert hb["q2_q4_wait"][ix("q2_skipped_after_removal")] >= 1

    # ---- searched: small random instances until every countable quirk has been seen at least twice
    TaskEnv = R.ref_taskenv_class()
    want = [k for k in CENSUS_KEYS if k not in ("decisions", "q10_max_groups", "q10_multi_group_slot", "q11_group_follows_to_depot")]
    cand = []
    for s in range(400):
        rs = np.random.default_rng(s)
        A, T, M = int(rs.integers(2, 6)), int(rs.integers(4, 10)), int(rs.integers(2, 4))
        env = TaskEnv((A, A), (T, T), 1, M, seed=s)
        env.max_waiting_time = 10
        ia = R.instance_arrays(env)
        ep, cen = record(env, "random", seed=5000 + s)
        if len(ep["leader"]) <= 120:
            cand.append((s, dict(ia, M=M), ep, cen))
    have = {k: int(total[ix(k)]) for k in want}
    for _ in range(12):
        need = [k for k in want if have[k] < 2]
        if not need:
            break
        best = max(cand, key=lambda c: (sum(min(c[3].c[k], 1) for k in need), -len(c[2]["leader"])))
        if sum(best[3].c[k] for k in need) == 0:
            break
        cand.remove(best)
        s, inst, ep, cen = best
        keep(f"seed{s}_{inst['A']}x{len(inst['req'])}", inst, ep, cen)
        total += cen.vector()
        for k in want:
            have[k] += cen.c[k]
    print("census over all fixtures:", dict(zip(CENSUS_KEYS, total.tolist())))
    missing = [k for k in want if total[ix(k)] == 0]
    assert not missing, f"no fixture exercises {missing}"
    np.savez_compressed(OUT / "quirks.npz", names=np.array(names), census_keys=np.array(CENSUS_KEYS), census_total=total, **pack(eps), **extra)
    print("wrote", OUT / "quirks.npz", sum(len(e["leader"]) for e in eps), "decisions in", len(eps), "fixtures")


if __name__ == "__main__":
    sys.exit(main())
