"""oracle/canon.py -- canonical state form, digests and comparison helpers (TEST INFRASTRUCTURE ONLY).

The reference (oracle/ref_shim.canonical_state), the C oracle (OracleEnv.export) and the CUDA path
(BatchedTaskEnv.export_state) are all reduced to the same dict of flat arrays; `state_digest` / `obs_digest`
hash them byte-exactly so that per-step golden vectors recorded from the real reference stay small.
"""
from __future__ import annotations

import hashlib

import numpy as np

MC_CANON = 8      # member slots in the canonical form (legal trajectories never exceed max_coalition_size <= 8)

_STATE_KEYS = (
    ("n_mem", np.int32), ("members", np.int32), ("mem_arr", np.float64), ("status", np.int32),
    ("feasible", np.uint8), ("finished", np.uint8), ("time_start", np.float64), ("time_finish", np.float64),
    ("n_aband_task", np.int32), ("node", np.int32), ("has_route", np.uint8), ("last_arrival", np.float64),
    ("next_decision", np.float64), ("travel_dist", np.float64), ("assigned", np.uint8), ("returned", np.uint8),
    ("n_aband_agent", np.int32))
DISCRETE_KEYS = ("n_mem", "members", "status", "feasible", "finished", "n_aband_task", "node", "has_route",
                 "assigned", "returned", "n_aband_agent")
FLOAT_KEYS = ("mem_arr", "time_start", "time_finish", "last_arrival", "next_decision", "travel_dist")


def _canon_f64(a):
    a = np.array(a, dtype=np.float64, copy=True)
    a[np.isnan(a)] = np.nan          # one NaN payload
    a[a == 0] = 0.0                  # -0.0 -> +0.0
    return a


def normalise(state: dict, MC: int = MC_CANON) -> dict:
    """Pad/crop member slots to MC, zero the unused slots, canonicalise floats."""
    out = {}
    n_mem = np.asarray(state["n_mem"], np.int32)
    T = n_mem.shape[0]
    mem = np.full((T, MC), -1, np.int32)
    arr = np.zeros((T, MC), np.float64)
    src_m = np.asarray(state["members"]).reshape(T, -1)
    src_a = np.asarray(state["mem_arr"]).reshape(T, -1)
    for j in range(T):
        n = int(n_mem[j])
        assert n <= MC, f"task {j} has {n} members > canonical capacity {MC}"
        mem[j, :n] = src_m[j, :n]
        arr[j, :n] = src_a[j, :n]
    for k, dt in _STATE_KEYS:
        if k == "members":
            out[k] = mem
        elif k == "mem_arr":
            out[k] = _canon_f64(arr)
        elif dt is np.float64:
            out[k] = _canon_f64(state[k])
        else:
            out[k] = np.ascontiguousarray(state[k], dtype=dt)
    out["now"] = float(state["now"])
    return out


def state_digest(state: dict) -> int:
    s = normalise(state)
    h = hashlib.sha256()
    for k, _ in _STATE_KEYS:
        h.update(np.ascontiguousarray(s[k]).tobytes())
    h.update(np.float64(s["now"] + 0.0).tobytes())
    return int.from_bytes(h.digest()[:8], "little")


def obs_digest(mask, agent_obs, task_obs) -> int:
    """Digest of what the policy sees: u8 mask [T+1], fp32 agent rows [A,6], fp32 task rows [T+1,5]."""
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(mask, np.uint8).tobytes())
    a = np.array(agent_obs, np.float32, copy=True)
    t = np.array(task_obs, np.float32, copy=True)
    a[a == 0] = 0.0
    t[t == 0] = 0.0
    h.update(a.tobytes())
    h.update(t.tobytes())
    return int.from_bytes(h.digest()[:8], "little")


def diff_states(a: dict, b: dict, rtol: float = 0.0) -> list[str]:
    """Human-readable list of differences (empty == equal).  rtol applies to FLOAT_KEYS only."""
    a, b = normalise(a), normalise(b)
    out = []
    for k in DISCRETE_KEYS:
        if not np.array_equal(a[k], b[k]):
            idx = np.argwhere(np.asarray(a[k]) != np.asarray(b[k]))[:4].tolist()
            out.append(f"{k}: differs at {idx}: {np.asarray(a[k])[tuple(np.array(idx).T)]} vs {np.asarray(b[k])[tuple(np.array(idx).T)]}")
    for k in FLOAT_KEYS:
        x, y = a[k], b[k]
        if rtol == 0.0:
            ok = np.array_equal(x, y, equal_nan=True)
        else:
            ok = np.allclose(x, y, rtol=rtol, atol=0.0, equal_nan=True)
        if not ok:
            bad = np.argwhere(~((x == y) | (np.isnan(x) & np.isnan(y))))[:4].tolist()
            out.append(f"{k}: differs at {bad}: {x[tuple(np.array(bad).T)]} vs {y[tuple(np.array(bad).T)]}")
    if not (a["now"] == b["now"] or (rtol and abs(a["now"] - b["now"]) <= rtol * abs(b["now"]))):
        out.append(f"now: {a['now']!r} vs {b['now']!r}")
    return out
