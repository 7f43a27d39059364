"""oracle/make_golden.py -- regenerate tests/golden/ from the REAL reference (run in the build container only:
`python -m oracle.make_golden`).  TEST INFRASTRUCTURE ONLY.

What it records (all produced by executing /root/reference/env/task_env.py unmodified through oracle/ref_shim.py):
  instances_20A_50T.npz  static data of the 50 bundled pickles testSet_20A_50T_CONDET/env_i.pkl
  ctasd.json             CTAS-D routes (env_i/results.yaml, read as baselines/CTAS-D.py:10-46 does), the reference's own
                         known-answer rows (CTAS-D_300s.csv) and the metrics the reference produces here
  traces_pickles.npz     50 instances x {random, greedy-nearest}: (leader, action, followers) of every decision, the clock,
                         the step reward, a digest of (mask, agent obs, task obs) and a digest of the full live state at
                         every decision, final metrics / finished flags / final state digest
  traces_sweep.npz       same for fresh TaskEnv(seed=s) instances of 10A/20T, 20A/50T, 30A/100T, 50A/200T
  full_dump.npz          instance 0, both policies: the undigested per-decision obs, masks and states
It also re-checks, on every distance the reference computes, that np.linalg.norm == sqrt(fma(dy,dy,dx*dx)).
"""
from __future__ import annotations

import contextlib
import csv
import io
import json
import sys
import warnings
from pathlib import Path

import numpy as np

from . import canon
from . import ref_shim as R

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"
MAXF = 8  # follower slots stored per decision


def _check_norm_formula(n=20000):
    rng = np.random.default_rng(7)
    a, b = rng.random((n, 2)), rng.random((n, 2))
    bad = 0
    for p, q in zip(a, b):
        d = p - q
        ref = float(np.linalg.norm(d))
        mine = float(np.sqrt(np.float64(R._fma(float(d[1]), float(d[1]), float(d[0]) * float(d[0])))))
        bad += ref != mine
    return bad


def record_episode(env, policy, seed, keep_full=False):
    digs_o, digs_s, full = [], [], []

    def on_dec(env_, leader, mask, ag, tk):
        st = R.canonical_state(env_)
        digs_o.append(canon.obs_digest(mask, ag.astype(np.float32), tk.astype(np.float32)))
        digs_s.append(canon.state_digest(st))
        if keep_full:
            full.append(dict(mask=mask.copy(), agent_obs=ag.astype(np.float32), task_obs=tk.astype(np.float32),
                             state=canon.normalise(st)))

    trace, reward, fin = R.run_reference_episode(env, policy, seed, on_decision=on_dec)
    met = R.reference_metrics(env, fin)
    n = len(trace["leader"])
    fol = np.full((n, MAXF), -1, np.int8)
    for k, f in enumerate(trace["followers"]):
        assert len(f) <= MAXF
        fol[k, :len(f)] = f
    ep = dict(
        leader=np.array(trace["leader"], np.int8), action=np.array(trace["action"], np.int16),
        nfol=np.array([len(f) for f in trace["followers"]], np.int8), followers=fol,
        now=np.array(trace["now"], np.float64), reward=np.array(trace["reward"], np.float64),
        dig_obs=np.array(digs_o, np.uint64), dig_state=np.array(digs_s, np.uint64),
        metrics=np.array([reward, met["success_rate"], met["makespan"], met["time_cost"], met["waiting_time"],
                          met["travel_dist"], met["efficiency"], float(n)], np.float64),
        finished=np.asarray(fin, np.uint8),
        final_digest=np.array([canon.state_digest(R.canonical_state(env))], np.uint64))
    return ep, full


def pack(episodes):
    """list of per-episode dicts -> flat arrays + offsets."""
    out = {}
    off = np.zeros(len(episodes) + 1, np.int64)
    for i, e in enumerate(episodes):
        off[i + 1] = off[i] + len(e["leader"])
    out["offsets"] = off
    for k in ("leader", "action", "nfol", "followers", "now", "reward", "dig_obs", "dig_state"):
        out[k] = np.concatenate([e[k] for e in episodes], axis=0)
    out["metrics"] = np.stack([e["metrics"] for e in episodes])
    out["final_digest"] = np.concatenate([e["final_digest"] for e in episodes])
    return out


def main():
    warnings.filterwarnings("ignore")
    assert R.available(), "reference tree missing"
    OUT.mkdir(parents=True, exist_ok=True)
    bad = _check_norm_formula()
    print(f"norm formula mismatches: {bad}/20000")
    assert bad == 0, "np.linalg.norm is not sqrt(fma(dy,dy,dx*dx)) on this host; golden vectors would not be portable"

    # ---- instances --------------------------------------------------------------------------------------
    inst = [R.instance_arrays(R.load_pickle(i)) for i in range(50)]
    np.savez_compressed(OUT / "instances_20A_50T.npz",
                        task_xy=np.stack([x["task_xy"] for x in inst]), depot_xy=np.stack([x["depot_xy"] for x in inst]),
                        req=np.stack([x["req"] for x in inst]), dur=np.stack([x["dur"] for x in inst]), A=np.int32(20))

    # ---- CTAS-D known-answer path ------------------------------------------------------------------------
    rows = list(csv.DictReader(open(R.TESTSET / "CTAS-D_300s.csv")))
    ct = []
    for i in range(50):
        env = R.load_pickle(i)
        routes = R.ctasd_routes(i)
        for a, r in routes.items():
            env.pre_set_route(list(r), a)
        with contextlib.redirect_stdout(io.StringIO()):
            env.execute_by_route("./", "CTAS-D", False)
        reward, fin = env.get_episode_reward(100)
        met = R.reference_metrics(env, fin)
        ct.append(dict(routes={str(a): r for a, r in routes.items()},
                       csv={k: float(rows[i][k]) for k in ("success_rate", "makespan", "waiting_time", "travel_dist", "efficiency")},
                       ref_here=dict(met, reward=reward), finished=[int(x) for x in fin],
                       final_digest=str(canon.state_digest(R.canonical_state(env)))))
    json.dump(ct, open(OUT / "ctasd.json", "w"))
    print("ctasd done")

    # ---- trace golden: pickles ---------------------------------------------------------------------------
    eps, names, fulls = [], [], {}
    for i in range(50):
        for policy in ("random", "greedy"):
            ep, full = record_episode(R.load_pickle(i), policy, seed=i, keep_full=(i == 0))
            eps.append(ep)
            names.append(f"{i}/{policy}")
            if full:
                fulls[policy] = (ep, full)
    np.savez_compressed(OUT / "traces_pickles.npz", names=np.array(names), finished=np.stack([e["finished"] for e in eps]),
                        **pack(eps))
    print("pickle traces done:", sum(len(e["leader"]) for e in eps), "decisions")

    fd = {}
    for policy, (ep, full) in fulls.items():
        fd[f"{policy}/mask"] = np.stack([f["mask"] for f in full])
        fd[f"{policy}/agent_obs"] = np.stack([f["agent_obs"] for f in full])
        fd[f"{policy}/task_obs"] = np.stack([f["task_obs"] for f in full])
        for k in full[0]["state"]:
            fd[f"{policy}/state/{k}"] = np.stack([np.asarray(f["state"][k]) for f in full])
    np.savez_compressed(OUT / "full_dump.npz", **fd)

    # ---- trace golden: shape sweep on fresh generator instances --------------------------------------------
    TaskEnv = R.ref_taskenv_class()
    eps, names, insts = [], [], {}
    for (A, T) in ((10, 20), (20, 50), (30, 100), (50, 200)):
        for s in range(3):
            for policy in ("random", "greedy"):
                env = TaskEnv((A, A), (T, T), 1, R.COALITION_SIZE, seed=s)    # worker.py:32 positional order
                ia = R.instance_arrays(env)
                ep, _ = record_episode(env, policy, seed=1000 + s)
                eps.append(ep)
                names.append(f"{A}x{T}/{s}/{policy}")
                for k in ("task_xy", "depot_xy", "req", "dur"):
                    insts[f"inst/{A}x{T}/{s}/{k}"] = ia[k]
    sw = pack(eps)
    for i, e in enumerate(eps):
        insts[f"finished/{names[i]}"] = e["finished"]
    np.savez_compressed(OUT / "traces_sweep.npz", names=np.array(names), **sw, **insts)
    print("sweep traces done:", sum(len(e["leader"]) for e in eps), "decisions")


if __name__ == "__main__":
    sys.exit(main())
