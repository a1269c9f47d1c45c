"""Tensor-side descriptions of the two environments on the planning path (no simulator).

The planner only reads a handful of scalars/arrays from ``env`` (SURVEY.md 8b "Env hand-off"):
``agent_dim, object_dyn_dim, object_stat_dim, nObj, action_space.low/high, observation_space.shape, goal`` and, for
the task cost, ``case, sparse, shaped_reward, threshold, buffer_threshold``.  ``TensorEnv`` carries exactly those
and can stand in for ``FetchPickAndPlaceConstruction`` (mbrl/environments/fpp_construction_env.py:17-125) or
``PlaygroundwGoals`` (mbrl/environments/playground_env_wgoals.py:12-72) wherever MuJoCo is not available
(benchmarks, tests).  A real reference env object works too: ``describe_env`` reads the same attributes from it.
"""
from types import SimpleNamespace

import numpy as np

from . import _lib


class TensorEnv:
    def __init__(self, *, name, kind, n_obj, agent_dim, dyn_dim, stat_dim, action_dim, goal_dim, case="Singletower",
                 sparse=False, shaped_reward=True, threshold=0.02, buffer_threshold=0.025, height_offset=0.42):
        self.name = name
        self.kind = kind
        self.nObj = self.num_blocks = self.num_objs = n_obj
        self.agent_dim, self.object_dyn_dim, self.object_stat_dim = agent_dim, dyn_dim, stat_dim
        self.case, self.sparse, self.shaped_reward = case, sparse, shaped_reward
        self.threshold, self.buffer_threshold = threshold, buffer_threshold
        self.height_offset = height_offset
        self.gripper_init = np.array([1.344, 0.749, 0.5319])  # fpp_construction_env.py:119
        if case == "Flip":  # :93-105
            self.coord_weights = np.array([1.0, 0.0, 0.0], np.float32)
            self.buffer_threshold = np.float32(1e-4)
            self.cost_thres = self.coord_weights * np.float32(0.087)
        elif case == "Slide":
            self.coord_weights = np.ones(3, np.float32)
            self.buffer_threshold = np.array([0.1, 0.1, height_offset], np.float32)
            self.cost_thres = np.array([0.1, 0.1, height_offset], np.float32)
        sd = agent_dim + n_obj * (dyn_dim + stat_dim)
        self.observation_space_size_preproc = sd
        self.goal_space_size = goal_dim
        self.observation_space = SimpleNamespace(shape=(sd + goal_dim,))
        self.action_space = SimpleNamespace(low=-np.ones(action_dim, np.float32), high=np.ones(action_dim, np.float32),
                                            shape=(action_dim,))
        self.goal_idx = np.arange(sd, sd + goal_dim)
        self.goal = np.zeros(goal_dim, np.float32)


def construction_env(num_blocks, case="Singletower", sparse=False, shaped_reward=True, height_offset=0.42):
    """fpp_construction_env.py:32-35 (dims), :59-68 (thresholds)."""
    if "tower" in case or case == "Pyramid":
        thr = 0.02
    elif case == "PickAndPlace":
        thr = 0.025
    elif case == "Slide":
        thr = 0.1
    elif case == "Flip":
        thr = 0.087 * 2
    else:
        thr = 0.05
    return TensorEnv(name="FetchPickAndPlaceConstruction", kind="construction", n_obj=num_blocks, agent_dim=10,
                     dyn_dim=12, stat_dim=0, action_dim=4, goal_dim=3 * num_blocks + 3, case=case, sparse=sparse,
                     shaped_reward=shaped_reward, threshold=thr, buffer_threshold=0.025, height_offset=height_offset)


def playground_env(num_objs):
    """playground_env_wgoals.py:24-27."""
    return TensorEnv(name="PlaygroundwGoals", kind="playground", n_obj=num_objs, agent_dim=4, dyn_dim=6, stat_dim=3,
                     action_dim=2, goal_dim=2 * num_objs, case="playground", threshold=0.23, buffer_threshold=0.23)


def robodesk_env(task="open_slide", reward="dense"):
    """mbrl/environments/robodesk_env.py:14-31 (dims), robodesk/robodesk.py:70-95 (8 bodies, 6 object types, 8 action dims)."""
    env = TensorEnv(name="Robodesk", kind="robodesk", n_obj=8, agent_dim=24, dyn_dim=13, stat_dim=6, action_dim=8, goal_dim=0,
                    case=task, sparse=reward == "sparse", shaped_reward=False, threshold=0.0, buffer_threshold=0.0)
    env.task, env.reward = task, reward
    return env


def _arr(x):
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.asarray(x, np.float32)


def describe_env(env):
    """Reads the hot-path attributes from a TensorEnv or from a reference env object."""
    kind = getattr(env, "kind", None)
    if kind is None:
        cls = type(env).__name__
        kind = ("construction" if "Construction" in cls else "playground" if "Playground" in cls else
                "robodesk" if cls == "Robodesk" else "unknown")
    A, D, S, N = int(env.agent_dim), int(env.object_dyn_dim), int(env.object_stat_dim), int(env.nObj)
    O = int(env.observation_space.shape[0])
    U = int(env.action_space.shape[0])
    sd = A + N * (D + S)
    d = dict(kind=kind, n_obj=N, agent_dim=A, dyn_dim=D, stat_dim=S, action_dim=U, goal_dim=O - sd, obs_dim=O,
             state_dim=sd, action_low=np.asarray(env.action_space.low, np.float32),
             action_high=np.asarray(env.action_space.high, np.float32))
    d["env_kind"] = {"construction": _lib.ENV_CONSTRUCTION, "playground": _lib.ENV_PLAYGROUND,
                     "robodesk": _lib.ENV_ROBODESK}.get(kind, _lib.ENV_NONE)
    case = getattr(env, "case", "")
    d["case"] = case
    task = getattr(env, "task", None)
    d["task_cost_supported"] = kind in ("construction", "playground") or (kind == "robodesk" and task in _lib.ROBODESK_TASKS)
    # any other environment: its own cost_fn is called on the materialised trajectories (SURVEY.md 8b "Env hand-off")
    d["external_cost"] = (not d["task_cost_supported"] or bool(getattr(env, "force_external_cost", False))) and \
        callable(getattr(env, "cost_fn", None))
    if d["external_cost"]:
        d["env_kind"] = _lib.ENV_EXTERNAL
    d["cost_case"] = (_lib.CASE_FLIP if case == "Flip" else _lib.CASE_SLIDE if case == "Slide" else
                      _lib.CASE_TOWER if ("tower" in case or case == "Pyramid") else _lib.CASE_PICKANDPLACE)
    d["sparse"] = int(bool(getattr(env, "sparse", False)))
    if kind == "robodesk" and d["task_cost_supported"]:  # robodesk_env.py:53-73: cost_case = task, sparse = (reward == "sparse")
        d["cost_case"] = _lib.ROBODESK_TASKS.index(task)
        d["sparse"] = int(getattr(env, "reward", "dense") == "sparse")
    d["shaped_reward"] = int(bool(getattr(env, "shaped_reward", False)))
    d["threshold"] = float(getattr(env, "threshold", 0.0))
    bt = _arr(getattr(env, "buffer_threshold", 0.025))
    d["buffer_threshold"] = float(bt.reshape(-1)[0]) if bt.size == 1 else 0.025
    # Flip / Slide (fpp_construction_env.py:93-105): tensors on a reference env, arrays on a TensorEnv
    d["coord_weights"] = np.broadcast_to(_arr(getattr(env, "coord_weights", np.ones(3))).reshape(-1), (3,)).astype(np.float32)
    d["buffer_xyz"] = np.broadcast_to(bt.reshape(-1), (3,)).astype(np.float32)
    d["cost_thres"] = np.broadcast_to(_arr(getattr(env, "cost_thres", np.zeros(3))).reshape(-1), (3,)).astype(np.float32)
    d["gripper_init"] = _arr(getattr(env, "gripper_init", np.zeros(3))).reshape(3).astype(np.float32)
    return d
