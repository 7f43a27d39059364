"""Instance I/O against the reference's own files for test instance 0 (recorded by oracle/make_yaml_golden.py)."""
from pathlib import Path

import numpy as np
import pytest
import yaml

from dcmrta_b200 import instance_io as io
from helpers import ctasd, pickle_instances

GOLD = Path(__file__).resolve().parent / "golden" / "planner_yaml_env0.npz"


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_planner_files_equal_the_reference_files(tmp_path, gold):
    """TestSetGenerator.py:41-116 -- same keys in the same order, the same numbers bit for bit, the small files byte for byte."""
    inst = pickle_instances()[0]
    io.write_planner_files(tmp_path, inst["task_xy"], inst["depot_xy"], inst["req"], inst["dur"], agents_num=inst["A"],
                           folder="testSet_20A_50T_CONDET", index=0)
    g = yaml.safe_load(open(tmp_path / "graph.yaml"))
    assert list(g) == ["vehicle0"]
    keys = list(g["vehicle0"])
    assert keys == list(gold["edge_keys"]) + list(gold["node_keys"])
    edges = np.array([g["vehicle0"][k] for k in gold["edge_keys"]], np.float64)
    assert np.array_equal(edges[:, [0, 1, 2, 4]], gold["edges"][:, [0, 1, 2, 4]])     # topology: identical
    # distances / travel times: the bundled files were written by an older CPython whose math.hypot was libm's; CPython >= 3.10
    # has its own algorithm, so the same expression (TestSetGenerator.py:37) differs in the last bit for ~1/3 of the edges
    np.testing.assert_allclose(edges[:, [3, 5]], gold["edges"][:, [3, 5]], rtol=4.5e-16, atol=0)
    import math
    xy = inst["task_xy"]
    a, b = int(edges[7, 0]), int(edges[7, 1])
    assert edges[7, 3] == math.hypot(xy[a][0] - xy[b][0], xy[a][1] - xy[b][1]) and edges[7, 5] == edges[7, 3] / 0.2
    assert np.array_equal(np.array([g["vehicle0"][k] for k in gold["node_keys"]]), gold["nodes"])
    for name in ("task_param", "vehicle_param", "planner_param"):
        assert open(tmp_path / f"{name}.yaml").read() == str(gold[name]), name


def test_det_planner_has_one_vehicle_block_per_agent():
    inst = pickle_instances()[0]
    d = io.planner_dicts(inst["task_xy"], inst["depot_xy"], inst["req"], inst["dur"], agents_num=3, planner="TEAMPLANNER_DET")
    T = len(inst["req"])
    assert list(d["graph"]) == ["vehicle0", "vehicle1", "vehicle2"] and d["planner_param"]["vehNumPerType"] == [1, 1, 1]
    assert d["graph"]["vehicle2"][f"edge{T * (T - 1)}"][:2] == [T + 2, 0]                 # depot-out node of vehicle 2
    assert d["graph"]["vehicle2"][f"edge{T * (T - 1) + 1}"][:2] == [0, T + 3 + 2]
    with pytest.raises(ValueError):
        io.planner_dicts(inst["task_xy"], inst["depot_xy"], inst["req"], inst["dur"], 3, planner="nope")


def test_routes_reader_matches_ctasd_py(tmp_path, gold):
    """baselines/CTAS-D.py:10-46 on a results.yaml rebuilt from the recorded node lists == the routes the golden CTAS-D run used."""
    nodes = yaml.safe_load(str(gold["routes"]))
    (tmp_path / "planner_param.yaml").write_text(str(gold["planner_param"]))
    yaml.safe_dump({"result": {"flagSuccess": 1}, "vehicle": {k: {"id": int(k[2:]), "node": v} for k, v in nodes.items()}},
                   open(tmp_path / "results.yaml", "w"))
    routes = io.read_planner_routes(tmp_path)
    want = {int(a): r for a, r in ctasd()[0]["routes"].items()}
    assert routes == want
    arr, n = io.routes_to_arrays([routes, None], agents_num=20)
    assert arr.shape[:2] == (2, 20) and arr.dtype == np.int32 and n[1].sum() == 0
    for a, r in want.items():
        assert arr[0, a, :n[0, a]].tolist() == r and not arr[0, a, n[0, a]:].any()
    (tmp_path / "results.yaml").write_text("result: {flagSuccess: 0}\n")
    assert io.read_planner_routes(tmp_path) is None


def test_metrics_csv_rows(tmp_path):
    """baselines/CTAS-D.py:62-96: NaN row below full success, time_cost = sum of start times, pandas layout."""
    g = ctasd()[0]
    m = [g["ref_here"]["reward"], g["ref_here"]["success_rate"], g["ref_here"]["makespan"], g["ref_here"]["time_cost"],
         g["ref_here"]["waiting_time"], g["ref_here"]["travel_dist"], g["ref_here"]["efficiency"], 0]
    row = io.planner_metrics_row(m, time_start=[1.0, 2.5, float("nan")])
    assert row["time_cost"] == 103.5 and row["makespan"] == g["ref_here"]["makespan"]
    for k, v in g["csv"].items():
        if k != "time_cost":
            assert row[k] == pytest.approx(v, rel=1e-14)
    bad = io.planner_metrics_row([0, 0.98, 50.0, 1, 1, 1, 1, 0])
    assert bad["success_rate"] == 0.98 and all(np.isnan(bad[c]) for c in io.METRIC_COLUMNS[1:])
    none = io.planner_metrics_row(None, found=False)
    assert none["success_rate"] == 0 and np.isnan(none["makespan"])
    df = io.write_metrics_csv(tmp_path / "CTAS-D.csv", [row, bad, none])
    text = open(tmp_path / "CTAS-D.csv").read().splitlines()
    assert text[0] == ",success_rate,makespan,time_cost,waiting_time,travel_dist,efficiency" and len(text) == 4 and len(df) == 3
