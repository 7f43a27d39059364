#!/usr/bin/env python
"""oracle/make_yaml_golden.py -- TEST INFRASTRUCTURE.  Records the reference's own planner files of test instance 0
(/root/reference/testSet_20A_50T_CONDET/env_0/*.yaml, written by the reference's TestSetGenerator.py) into
tests/golden/planner_yaml_env0.npz: the graph as arrays (edge keys in file order, the 6 numbers of each edge, node durations),
the three small files verbatim, and the routes baselines/CTAS-D.py reads out of results.yaml."""
from pathlib import Path

import numpy as np
import yaml

REF = Path("/root/reference/testSet_20A_50T_CONDET/env_0")
ROOT = Path(__file__).resolve().parent.parent
g = yaml.safe_load(open(REF / "graph.yaml"))
assert list(g) == ["vehicle0"]
keys = list(g["vehicle0"])
edges = [k for k in keys if k.startswith("edge")]
nodes = [k for k in keys if k.startswith("node")]
out = dict(edge_keys=np.array(edges), edges=np.array([g["vehicle0"][k] for k in edges], np.float64),
           node_keys=np.array(nodes), nodes=np.array([g["vehicle0"][k] for k in nodes], np.float64),
           key_order_ok=np.array(keys == edges + nodes))
for name in ("task_param", "vehicle_param", "planner_param"):
    out[name] = np.array(open(REF / f"{name}.yaml").read())
res = yaml.safe_load(open(REF / "results.yaml"))
out["routes"] = np.array(yaml.safe_dump({k: v["node"] for k, v in res["vehicle"].items()}))
np.savez_compressed(ROOT / "tests" / "golden" / "planner_yaml_env0.npz", **out)
print(len(edges), "edges", len(nodes), "nodes")
