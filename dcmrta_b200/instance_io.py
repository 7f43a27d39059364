"""dcmrta_b200/instance_io.py -- the disk formats either side of the step path.

  write_planner_files   the four YAML files `TestSetGenerator.py:41-116` writes next to every test instance for the CTAS-D
                        planner (vehicle_param, task_param, planner_param, graph: edges [from, to, 0, d, 0, d / 0.2])
  read_planner_routes   `baselines/CTAS-D.py:10-46`: per-agent action lists out of the planner's results.yaml
  routes_to_arrays      pad those lists into the [B, A, L] int32 tensor `dcm_execute_by_route` takes
  planner_metrics_row   the row `baselines/CTAS-D.py:73-94` appends to <method>.csv for one executed instance
  write_metrics_csv     the CSV itself (same columns / index as pandas.DataFrame(perf_metrics).to_csv)
Reference pickles are imported by the facade (dcmrta_b200/task_env.py, RL_test.py:34-44); instances are generated on the
device by dcm_generate (task_env.py:57-114 distributions).
"""
from __future__ import annotations

import math
from itertools import permutations
from pathlib import Path

import numpy as np

VELOCITY = 0.2          # task_env.py:99


def planner_dicts(task_xy, depot_xy, req, dur, agents_num: int, planner: str = "TEAMPLANNER_CONDET", solver_time: float = 300.0,
                  folder: str = "testSet", index: int = 0):
    """The four dictionaries of TestSetGenerator.py:41-112 for one instance, with plain Python scalars so that yaml.dump
    renders them exactly as the reference does."""
    T = len(task_xy)
    coords = [(float(x), float(y)) for x, y in np.asarray(task_xy, np.float64)]
    depot = (float(depot_xy[0]), float(depot_xy[1]))
    dist = lambda a, b: math.hypot(a[0] - b[0], a[1] - b[1])                   # TestSetGenerator.py:37
    depot_distance = [dist(depot, c) for c in coords]                           # :51
    pairs = list(permutations(range(T), 2))                                     # :52

    def vehicle_graph(start_node, end_node):
        g = {f"edge{i}": [a, b, 0, dist(coords[a], coords[b]), 0, float(dist(coords[a], coords[b]) / VELOCITY)] for i, (a, b) in enumerate(pairs)}
        for j in range(T):                                                      # :59-61 depot -> task, task -> depot
            g[f"edge{2 * j + len(pairs)}"] = [start_node, j, 0, depot_distance[j], 0, depot_distance[j] / VELOCITY]
            g[f"edge{2 * j + len(pairs) + 1}"] = [j, end_node, 0, depot_distance[j], 0, depot_distance[j] / VELOCITY]
        for j in range(T):                                                      # :62-63
            g[f"node{j}"] = float(dur[j])
        return g

    agent_yaml, graph_yaml = {}, {}
    if planner == "TEAMPLANNER_CONDET":
        agent_yaml["vehicle0"] = {"engCap": 1e6, "engCost": 0., "capVector": [1.0], "capVar": [0.]}
        graph_yaml["vehicle0"] = vehicle_graph(T, T + 1)
    elif planner == "TEAMPLANNER_DET":
        for a in range(agents_num):                                             # :64-75
            agent_yaml[f"vehicle{a}"] = {"engCap": 1e6, "engCost": 1., "capVector": [1.0], "capVar": [0.]}
            graph_yaml[f"vehicle{a}"] = vehicle_graph(T + a, T + agents_num + a)
    else:
        raise ValueError(f"unknown planner {planner!r}")
    task_yaml = {f"task{j}": {"and0": {"or0": {"geq": True, "capId": 0, "capReq": float(req[j]), "capVar": 0.}}} for j in range(T)}   # :77-78
    base = f"./{folder}/env_{index}"
    planner_param = {                                                           # :83-113
        "flagOptimizeCost": True, "flagTaskComplete": True, "flagSprAddCutToSameType": True, "taskCompleteReward": 10000,
        "timePenalty": 100, "recoursePenalty": 1.0, "taskRiskPenalty": 0.0, "LARGETIME": 10000.0, "MAXTIME": 1000.0, "MAXENG": 1E8,
        "flagSolver": planner, "CcpBeta": 0.95, "taskBeta": 0.95, "solverMaxTime": solver_time, "solverIterMaxTime": 50.0,
        "flagNotUseUnralavant": True, "MAXALPHA": 20.0, "taskNum": int(T),
        "vehNum": 1 if planner == "TEAMPLANNER_CONDET" else int(agents_num), "capNum": 1, "vehTypeNum": 1,
        "vehNumPerType": [int(agents_num)] if planner == "TEAMPLANNER_CONDET" else [1] * int(agents_num),
        "sampleNum": 500, "randomType": 0, "capType": [0],
        "vehicleParamFile": f"{base}/vehicle_param.yaml", "taskParamFile": f"{base}/task_param.yaml", "graphFile": f"{base}/graph.yaml",
    }
    return {"vehicle_param": agent_yaml, "task_param": task_yaml, "planner_param": planner_param, "graph": graph_yaml}


def write_planner_files(out_dir, task_xy, depot_xy, req, dur, agents_num, **kw):
    """TestSetGenerator.py:79-116: <out_dir>/{vehicle_param,task_param,planner_param,graph}.yaml"""
    import yaml
    out = Path(out_dir)
    out.mkdir(parents=True, exist_ok=True)
    d = planner_dicts(task_xy, depot_xy, req, dur, agents_num, **kw)
    for name, content in d.items():
        with open(out / f"{name}.yaml", "w") as f:
            yaml.dump(content, f, sort_keys=False)
    return d


def read_planner_routes(env_dir):
    """baselines/CTAS-D.py:10-46: {agent: [actions]} from <env_dir>/results.yaml (node lists with the first node dropped,
    vehicles whose list is [0] skipped; 0 = depot, j + 1 = task j), or None when the planner found nothing."""
    import yaml
    d = Path(env_dir)
    with open(d / "planner_param.yaml") as f:
        p = yaml.safe_load(f)
    num_veh = p["vehNum"] if p["flagSolver"] == "TEAMPLANNER_DET" else p["vehNumPerType"][0]
    if not (d / "results.yaml").exists():
        return None
    with open(d / "results.yaml") as f:
        data = yaml.safe_load(f)
    if "vehicle" not in data:
        return None
    nodes = [data["vehicle"][f"vv{v + 1}"]["node"] for v in range(num_veh) if f"vv{v + 1}" in data["vehicle"]]
    return {a: list(r)[1:] for a, r in enumerate(nodes) if r != [0]}


def routes_to_arrays(routes_per_env, agents_num: int):
    """[{agent: [actions]}] -> (routes [B, A, L] int32 zero padded, route_len [B, A] int32) for BatchedTaskEnv.execute_by_route"""
    B = len(routes_per_env)
    L = max([len(r) for rs in routes_per_env for r in (rs or {}).values()] + [1])
    routes = np.zeros((B, agents_num, L), np.int32)
    rlen = np.zeros((B, agents_num), np.int32)
    for b, rs in enumerate(routes_per_env):
        for a, r in (rs or {}).items():
            routes[b, int(a), :len(r)] = r
            rlen[b, int(a)] = len(r)
    return routes, rlen


METRIC_COLUMNS = ("success_rate", "makespan", "time_cost", "waiting_time", "travel_dist", "efficiency")


def planner_metrics_row(metrics_row, time_start=None, found=True):
    """baselines/CTAS-D.py:62-94.  metrics_row: one row of dcm_compute_metrics [reward, success_rate, makespan, time_cost,
    waiting_time, travel_dist, efficiency, decisions]; time_start: the env's time_start vector (0 where not feasible) -- the
    CSV's time_cost is the SUM of the start times (:88), not worker.py's mean."""
    nan = float("nan")
    if not found:
        return dict(zip(METRIC_COLUMNS, (0, nan, nan, nan, nan, nan)))
    sr = float(metrics_row[1])
    if sr < 1:
        return dict(zip(METRIC_COLUMNS, (sr, nan, nan, nan, nan, nan)))
    tc = float(np.sum(np.nan_to_num(np.asarray(time_start, np.float64), nan=100))) if time_start is not None else float(metrics_row[3])
    return dict(zip(METRIC_COLUMNS, (sr, float(metrics_row[2]), tc, float(metrics_row[4]), float(metrics_row[5]), float(metrics_row[6]))))


def write_metrics_csv(path, rows):
    """pandas.DataFrame(perf_metrics).to_csv(path) of baselines/CTAS-D.py:96: index column + the six metrics"""
    import pandas as pd
    df = pd.DataFrame({c: [r[c] for r in rows] for c in METRIC_COLUMNS})
    df.to_csv(path)
    return df
