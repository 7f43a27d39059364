"""GPU tests of the TaskEnv-compatible facade: the reference's own caller loops (worker.py:45-87, baselines/CTAS-D.py:60-94)
run against dcmrta_b200.task_env.TaskEnv and must reproduce the golden vectors recorded from the real reference."""
import copy
import pickle

import numpy as np
import pytest

from oracle import canon

from helpers import ctasd, pickle_instances, pickle_traces

pytestmark = pytest.mark.gpu


def as_dicts(inst):
    """(task_dic, agent_dic, depot) in the reference's dict-of-dicts form (what RL_test.py:36-41 passes to reset)."""
    T = inst["task_xy"].shape[0]
    tasks = {j: dict(ID=j, location=inst["task_xy"][j], requirements=np.array([inst["req"][j]]), time=np.array([inst["dur"][j]])) for j in range(T)}
    agents = {i: dict(ID=i) for i in range(inst["A"])}
    return tasks, agents, dict(location=inst["depot_xy"], members=[])


def worker_loop(env, ep, tr, check=True):
    """worker.py:45-87 with the policy / leader / followers replaced by the recorded trace."""
    k = 0
    while not env.finished and env.current_time < 100:
        ids, t = env.next_decision()
        groups = env.get_unique_group(ids)
        env.current_time = t
        env.task_update()
        env.agent_update()
        for group in groups:
            while len(group) > 0:
                leader = int(ep["leader"][k])
                assert leader in group
                agent = env.agent_dic[leader]
                assert not agent["returned"]
                mask = env.get_unfinished_task_mask()
                mask = np.insert(mask, 0, False) if np.sum(mask) == env.tasks_num else np.insert(mask, 0, True)
                ag = np.float32(env.get_current_agent_status(agent))
                tk = np.float32(env.get_current_task_status(agent))
                if check:
                    assert env.current_time == ep["now"][k]
                    assert canon.obs_digest(mask.astype(np.uint8), ag, tk) == int(ep["dig_obs"][k]), k
                fol = tr.followers(ep, k)
                env.random_choice = lambda a, size=None, replace=True, _f=fol: np.array(_f, dtype=np.int64)
                group, r = env.step(group, leader, int(ep["action"][k]), k)
                if check:
                    assert r == ep["reward"][k]
                env.task_update()
                env.agent_update()
                k += 1
        env.finished = env.check_finished()
    return k


def perf_metrics(env, finished_tasks):
    """worker.py:103-108"""
    return dict(success_rate=np.sum(finished_tasks) / len(finished_tasks), makespan=env.current_time,
                time_cost=np.nanmean(env.get_matrix(env.task_dic, "time_start")),
                waiting_time=np.mean(env.get_matrix(env.agent_dic, "sum_waiting_time")),
                travel_dist=np.sum(env.get_matrix(env.agent_dic, "travel_dist")),
                efficiency=np.mean(env.get_matrix(env.task_dic, "sum_waiting_time")))


@pytest.mark.parametrize("episode", [0, 1, 4, 9, 12, 17, 20, 23, 28, 33, 37, 40, 45, 48, 53, 56, 61, 66, 69, 74, 79, 82, 87, 90, 95, 99])
def test_worker_loop_through_facade(episode):
    from dcmrta_b200.task_env import TaskEnv
    tr = pickle_traces()
    ep = tr.episode(episode)
    inst = pickle_instances()[int(ep["name"].split("/")[0])]
    env = TaskEnv((20, 20), (50, 50), 1, 5, seed=0)
    env.max_waiting_time = 10                       # RL_test.py:39-43
    env.reactive_planning = False
    env.reset(as_dicts(inst))
    env.clear_decisions()
    n = worker_loop(env, ep, tr)
    assert n == len(ep["leader"])
    reward, fin = env.get_episode_reward(100)
    m = perf_metrics(env, fin)
    gold = ep["metrics"]
    assert reward == gold[0]
    for i, key in enumerate(("success_rate", "makespan", "time_cost", "waiting_time", "travel_dist", "efficiency")):
        assert m[key] == pytest.approx(gold[1 + i], rel=1e-12, abs=0), key
    assert np.array_equal(np.array(fin, bool), tr.finished[episode].astype(bool))
    # routes kept on the host mirror the decisions
    assert sum(len(r) for r in env.get_matrix(env.agent_dic, "route")) >= n


def test_ctasd_through_facade():
    """baselines/CTAS-D.py:60-94 against the reference's CTAS-D_300s.csv"""
    from dcmrta_b200.task_env import TaskEnv
    inst = pickle_instances()
    gold = ctasd()
    env = TaskEnv((20, 20), (50, 50), 1, 5, seed=0)
    for i in (0, 7, 23):
        env.reactive_planning = False
        env.reset(as_dicts(inst[i]))
        env.clear_decisions()
        for a, r in gold[i]["routes"].items():
            env.pre_set_route(copy.copy(r), int(a))
        env.force_wait = True
        env.execute_by_route("./", "CTAS-D", False)
        reward, fin = env.get_episode_reward(100)
        assert np.sum(fin) / len(fin) == gold[i]["csv"]["success_rate"]
        assert env.current_time == pytest.approx(gold[i]["csv"]["makespan"], rel=1e-14)
        assert np.mean(env.get_matrix(env.agent_dic, "sum_waiting_time")) == pytest.approx(gold[i]["csv"]["waiting_time"], rel=1e-12)
        assert np.sum(env.get_matrix(env.agent_dic, "travel_dist")) == pytest.approx(gold[i]["csv"]["travel_dist"], rel=1e-14)
        assert np.mean(env.get_matrix(env.task_dic, "sum_waiting_time")) == pytest.approx(gold[i]["csv"]["efficiency"], rel=1e-12)


def test_deepcopy_and_pickle_are_independent_snapshots():
    from dcmrta_b200.task_env import TaskEnv
    tr = pickle_traces()
    ep = tr.episode(4)
    inst = pickle_instances()[int(ep["name"].split("/")[0])]
    env = TaskEnv((20, 20), (50, 50), 1, 5, seed=0)
    env.reset(as_dicts(inst))
    env.clear_decisions()
    snap = copy.deepcopy(env)                        # worker.py:33 baseline_env = copy.deepcopy(env)
    blob = pickle.dumps(env)                         # TestSetGenerator.py:18
    worker_loop(env, ep, tr)
    r1, _ = env.get_episode_reward(100)
    assert snap.current_time == 0 and not any(snap.get_matrix(snap.agent_dic, "route"))
    worker_loop(snap, ep, tr)
    r2, _ = snap.get_episode_reward(100)
    env3 = pickle.loads(blob)
    worker_loop(env3, ep, tr)
    r3, _ = env3.get_episode_reward(100)
    assert r1 == r2 == r3 == ep["metrics"][0]


def test_individual_agent_loop_through_facade():
    """worker.py:159-198 `run_test_IS` (RL_test.py METHOD = 'IA'): every decider acts on its own through agent_step, in id order, no groups,
    no followers -- recorded from the real reference with the greedy-nearest policy (oracle/make_facade_golden.py)."""
    from pathlib import Path
    from dcmrta_b200.task_env import TaskEnv
    z = np.load(Path(__file__).resolve().parent / "golden" / "facade_is.npz")
    inst = pickle_instances()
    for i in (int(x) for x in z["instances"]):
        env = TaskEnv((20, 20), (50, 50), 1, 5, seed=0)
        env.max_waiting_time = 10
        env.reactive_planning = False
        env.reset(as_dicts(inst[i]))
        env.clear_decisions()
        k = 0
        while not env.finished and env.current_time < 100:                            # worker.py:163
            decision_agents, current_time = env.next_decision()
            env.current_time = current_time
            env.task_update()
            env.agent_update()
            for agent_id in decision_agents:                                          # worker.py:170
                agent = env.agent_dic[agent_id]
                if not agent["returned"]:
                    mask = env.get_unfinished_task_mask()
                    mask = np.insert(mask, 0, False) if np.sum(mask) == env.tasks_num else np.insert(mask, 0, True)
                    ag = np.float32(env.get_current_agent_status(agent))
                    tk64 = np.asarray(env.get_current_task_status(agent), np.float64)
                    assert int(agent_id) == int(z[f"{i}/agent"][k]) and env.current_time == z[f"{i}/now"][k], (i, k)
                    assert canon.obs_digest(mask.astype(np.uint8), ag, np.float32(tk64)) == int(z[f"{i}/dig_obs"][k]), (i, k)
                    action = int(z[f"{i}/action"][k])
                    env.agent_step(int(agent_id), action)                             # worker.py:186
                    env.task_update()
                    env.agent_update()
                    k += 1
            env.finished = env.check_finished()
        assert k == len(z[f"{i}/agent"])
        reward, fin = env.get_episode_reward(100)
        m = perf_metrics(env, fin)
        gold = z[f"{i}/metrics"]
        assert reward == gold[0] and np.array_equal(np.array(fin, np.uint8), z[f"{i}/finished"])
        for c, key in enumerate(("success_rate", "makespan", "time_cost", "waiting_time", "travel_dist", "efficiency")):
            assert m[key] == pytest.approx(gold[1 + c], rel=1e-12, abs=0), (i, key)


class _OldSchemaEnv:
    """Stand-in for the OLDER TaskEnv class the 50 bundled pickles were written by (SURVEY 8(c) "Pickle schema caveat")."""


def test_reference_schema_pickle_loads_through_setstate():
    """RL_test.py:35-44 with the class swapped: a pickle in the schema of the bundled env_i.pkl -- attributes coalition_size / tasks /
    agents / cost, max_waiting_time = 3, task keys start_time / finish_time / members_tmp ... and none of feasible_assignment /
    time_start / abandoned_agent, agent keys occupied / doing_task and no arrival_time / assigned, a depot without 'ID', per-task `time`
    as ndarray(1,) -- is unpickled into dcmrta_b200.task_env.TaskEnv, normalised as RL_test.py does, and replays a recorded episode."""
    import io
    from dcmrta_b200.task_env import TaskEnv
    tr = pickle_traces()
    ep = tr.episode(12)
    i = int(ep["name"].split("/")[0])
    inst = pickle_instances()[i]
    T, A = inst["task_xy"].shape[0], inst["A"]
    old = _OldSchemaEnv()
    old.coalition_size, old.tasks, old.agents, old.tasks_temp, old.agents_temp, old.cost = 5, (50, 50), (20, 20), T, A, None
    old.tasks_num, old.agents_num, old.traits_dim, old.max_waiting_time, old.current_time, old.dt, old.finished = T, A, 1, 3, 0, 0.1, False
    old.task_dic = {j: dict(ID=j, requirements=np.array([inst["req"][j]]), members=[], members_tmp=[], cost=[], location=inst["task_xy"][j],
                            finished=False, start_time=0, finish_time=0, status=np.array([inst["req"][j]]), time=np.array([inst["dur"][j]]),
                            label=0, consumed_time=0, start_tick=0, finish_tick=0) for j in range(T)}
    old.agent_dic = {a: dict(ID=a, abilities=np.ones(1), location=inst["depot_xy"], route=[], current_task=-1, contributed=False,
                             travel_time=0, velocity=0.2, next_decision=0, depot=inst["depot_xy"], travel_dist=0, occupied=False,
                             doing_task=False, current_waiting_time=0, cost=np.zeros(1)) for a in range(A)}
    old.depot = dict(location=inst["depot_xy"], members=list(range(A)))
    blob = pickle.dumps(old)

    class Unpickler(pickle.Unpickler):                       # the drop-in: wherever the pickle says TaskEnv, hand out ours
        def find_class(self, module, name):
            return TaskEnv if name == "_OldSchemaEnv" else super().find_class(module, name)
    env = Unpickler(io.BytesIO(blob)).load()
    assert isinstance(env, TaskEnv) and env.tasks_num == T and env.agents_num == A
    agents, tasks, depot = env.agent_dic, env.task_dic, env.depot                      # RL_test.py:36-43
    env.max_waiting_time = 10
    env.reactive_planning = False
    env.reset((tasks, agents, depot))
    env.clear_decisions()
    env.force_waiting = True
    assert env.max_waiting_time == 10 and float(tasks[3]["time"]) == inst["dur"][3]
    n = worker_loop(env, ep, tr)
    assert n == len(ep["leader"])
    reward, fin = env.get_episode_reward(100)
    assert reward == ep["metrics"][0]
