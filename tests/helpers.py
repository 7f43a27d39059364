"""Shared test helpers: golden loaders and trace replay drivers (tests only)."""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np

GOLD = Path(__file__).resolve().parent / "golden"
METRIC_KEYS = ("reward", "success_rate", "makespan", "time_cost", "waiting_time", "travel_dist", "efficiency", "n_steps")


def pickle_instances():
    z = np.load(GOLD / "instances_20A_50T.npz")
    return [dict(A=int(z["A"]), task_xy=z["task_xy"][i], depot_xy=z["depot_xy"][i], req=z["req"][i], dur=z["dur"][i])
            for i in range(z["task_xy"].shape[0])]


class Traces:
    """Flat trace arrays + offsets -> per-episode views."""

    def __init__(self, path):
        self.z = np.load(path)
        self.names = [str(n) for n in self.z["names"]]
        self.off = self.z["offsets"]

    def __len__(self):
        return len(self.names)

    def episode(self, i):
        s, e = int(self.off[i]), int(self.off[i + 1])
        ep = {k: self.z[k][s:e] for k in ("leader", "action", "nfol", "followers", "now", "reward", "dig_obs", "dig_state")}
        ep["metrics"] = self.z["metrics"][i]
        ep["final_digest"] = int(self.z["final_digest"][i])
        ep["name"] = self.names[i]
        return ep

    def followers(self, ep, k):
        return [int(x) for x in ep["followers"][k, :int(ep["nfol"][k])]]


def pickle_traces():
    t = Traces(GOLD / "traces_pickles.npz")
    t.finished = t.z["finished"]
    return t


def sweep_traces():
    return Traces(GOLD / "traces_sweep.npz")


def sweep_instance(tr: Traces, name: str):
    shape, s, _ = name.split("/")
    A, T = (int(x) for x in shape.split("x"))
    g = lambda k: tr.z[f"inst/{shape}/{s}/{k}"]
    return dict(A=A, task_xy=g("task_xy"), depot_xy=g("depot_xy"), req=g("req"), dur=g("dur"))


def ctasd():
    return json.load(open(GOLD / "ctasd.json"))


def full_dump():
    return np.load(GOLD / "full_dump.npz")


def quirk_traces():
    """tests/golden/quirks.npz (oracle/make_quirk_golden.py): hand-built and searched small fixtures, one quirk of SURVEY App. A each."""
    t = Traces(GOLD / "quirks.npz")
    t.census_keys = [str(k) for k in t.z["census_keys"]]
    return t


def quirk_instance(tr: Traces, name: str):
    g = lambda k: tr.z[f"inst/{name}/{k}"]
    return dict(A=int(g("A")), task_xy=g("task_xy"), depot_xy=g("depot_xy"), req=g("req").astype(np.int32), dur=g("dur"))


def quirk_census(tr: Traces, name: str):
    return dict(zip(tr.census_keys, (int(x) for x in tr.z[f"census/{name}"])))
