"""Builds the reference's own (unmodified) model / controller / env objects without MuJoCo.

TEST / BASELINE INFRASTRUCTURE ONLY (see ref_shims.py).  Never imported by product code.  The reference package is taken
from ``/root/reference`` where that is mounted (the build container: ``tests/golden/make_golden.py`` generates the fixtures
there) and otherwise from ``oracle/_ref/reference_mbrl.zip`` -- an archive of the reference's own, unmodified ``mbrl/*.py``
that ``oracle/build_ref.py`` (called by ``__graft_entry__.build()``) packs in the build container; it is git-ignored, travels
to the GPU box with the snapshot and is imported through ``zipimport``.  ``bench.py --impl reference`` times the reference's
``TorchMpcICem`` / ``TorchCuriosityMpcICem`` + ``GNNForwardEnsembleModel`` / ``ParallelNNDeterministicEnsemble`` from it.

Environment objects are created with ``object.__new__`` and given exactly the attributes the hot path
reads (fpp_construction_env.py:32-125, playground_env_wgoals.py:24-72); every *method* that runs
(``cost_fn``, ``obs_preproc``, ``obs_postproc``, the ``*_tensor`` helpers) is the reference's code.
"""
import os
import sys
import tempfile
import types
import warnings

import numpy as np
import torch

REF_ZIP = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "reference_mbrl.zip")
REFERENCE_ROOT = os.environ.get("CEEUS_REFERENCE_ROOT") or (
    "/root/reference" if os.path.isdir("/root/reference/mbrl") else REF_ZIP)


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "mbrl")) or (
        REFERENCE_ROOT.endswith(".zip") and os.path.isfile(REFERENCE_ROOT))


def read_reference_text(relpath):
    """Text of a file of the reference checkout (path relative to its root), from the tree or from the archive."""
    if os.path.isdir(REFERENCE_ROOT):
        return open(os.path.join(REFERENCE_ROOT, relpath)).read()
    import zipfile

    with zipfile.ZipFile(REFERENCE_ROOT) as z:
        return z.read(relpath).decode()


def reference_has(relpath):
    if os.path.isdir(REFERENCE_ROOT):
        return os.path.exists(os.path.join(REFERENCE_ROOT, relpath))
    import zipfile

    with zipfile.ZipFile(REFERENCE_ROOT) as z:
        return relpath in z.namelist()


_ready = False


def _prepare():
    global _ready
    if _ready:
        return
    if not reference_available():
        raise RuntimeError("reference package not found (neither a tree nor oracle/_ref/reference_mbrl.zip): %s" % REFERENCE_ROOT)
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)
    import ref_shims  # noqa: F401

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from mbrl import allogger

    logdir = tempfile.mkdtemp(prefix="ceeus_allogger_")
    allogger.basic_configure(logdir=logdir, default_outputs=["tensorboard"], manual_flush=True)
    warnings.filterwarnings("ignore", message=".*resized.*")
    warnings.filterwarnings("ignore", category=UserWarning)
    _ready = True


def make_construction_env(num_blocks, case="Singletower", sparse=False, shaped_reward=True, goal=None):
    _prepare()
    from gym import spaces
    from mbrl.environments.fpp_construction_env import FetchPickAndPlaceConstruction

    N = num_blocks
    env = object.__new__(FetchPickAndPlaceConstruction)
    env.name = "FetchPickAndPlaceConstruction"
    env.num_blocks = N
    env.case = case
    env.shaped_reward = shaped_reward
    env.sparse = sparse
    env.agent_dim, env.object_dyn_dim, env.object_stat_dim, env.nObj = 10, 12, 0, N
    orig = 10 + 12 * N
    gsz = 3 * N + 3
    env.goal_idx = np.arange(orig, orig + gsz)
    if case == "Flip":
        ach = [np.arange(10 + i * 12 + 3, 10 + i * 12 + 6) for i in range(N)]
    else:
        ach = [np.arange(10 + i * 12, 10 + i * 12 + 3) for i in range(N)]
    ach.append([0, 1, 2])
    env.achieved_goal_idx = np.asarray(ach).flatten()
    env.observation_space = spaces.Box(-np.inf, np.inf, shape=(orig + gsz,), dtype="float32")
    env.action_space = spaces.Box(-1.0, 1.0, shape=(4,), dtype="float32")
    env.observation_space_size_preproc = orig
    env.goal_space_size = gsz
    env.height_offset = 0.42
    if "tower" in case or case == "Pyramid":
        env.threshold = 0.02
    elif case == "PickAndPlace":
        env.threshold = 0.025
    elif case == "Slide":
        env.threshold = 0.1
    elif case == "Flip":
        env.threshold = 0.087 * 2
    else:
        env.threshold = 0.05
    env.goal_idx_tensor = torch.tensor(env.goal_idx, dtype=torch.int32)
    env.achieved_goal_idx_tensor = torch.tensor(env.achieved_goal_idx, dtype=torch.int32)
    if case == "Flip":
        env.coord_weights = torch.tensor([[1.0, 0.0, 0.0]], dtype=torch.float32)
        env.buffer_threshold = torch.tensor(1e-4, dtype=torch.float32)
        env.cost_thres = env.coord_weights * 0.087
    elif case == "Slide":
        env.coord_weights = torch.tensor([1.0, 1.0, 1.0], dtype=torch.float32)
        env.buffer_threshold = torch.tensor([0.1, 0.1, env.height_offset], dtype=torch.float32)
        env.cost_thres = env.coord_weights * 0.1
        env.cost_thres[-1] = env.height_offset
    else:
        env.buffer_threshold = torch.tensor(0.025, dtype=torch.float32)
    env.robot_base_xy = np.array([0.69189994, 0.74409998])
    env.manipulability_r = 0.8531
    env.robot_base_xy_tensor = torch.tensor(env.robot_base_xy, dtype=torch.float32)
    env.gripper_init = np.array([1.344, 0.749, 0.5319])
    env.gripper_init_tensor = torch.tensor(env.gripper_init, dtype=torch.float32)
    env.goal = np.zeros(gsz, dtype=np.float32) if goal is None else np.asarray(goal, dtype=np.float32)
    return env


def make_playground_env(num_objs, goal=None):
    _prepare()
    from gym import spaces
    from mbrl.environments.playground_env_wgoals import PlaygroundwGoals

    N = num_objs
    env = object.__new__(PlaygroundwGoals)
    env.name = "PlaygroundwGoals"
    env.num_objs = env.nObj = N
    env.dofs = 2
    env.agent_dim, env.object_dyn_dim, env.object_stat_dim = 4, 6, 3
    env.model = types.SimpleNamespace(nq=2 + 3 * N, nv=2 + 3 * N)
    orig = 4 + 9 * N
    gsz = 2 * N
    env.goal_idx = np.arange(orig, orig + gsz)
    env.achieved_goal_idx = np.asarray([np.arange(4 + 6 * i, 4 + 6 * i + 2) for i in range(N)]).flatten()
    env.observation_space = spaces.Box(-np.inf, np.inf, shape=(orig + gsz,), dtype="float32")
    env.action_space = spaces.Box(-1.0, 1.0, shape=(2,), dtype="float32")
    env.observation_space_size_preproc = orig
    env.goal_space_size = gsz
    env.goal_idx_tensor = torch.tensor(env.goal_idx, dtype=torch.int32)
    env.achieved_goal_idx_tensor = torch.tensor(env.achieved_goal_idx, dtype=torch.int32)
    env.goal = np.zeros(gsz, dtype=np.float32) if goal is None else np.asarray(goal, dtype=np.float32)
    return env


def make_robodesk_env(task="open_slide", reward="dense"):
    """Robodesk without MuJoCo: the attributes cost_fn reads (robodesk_env.py:14-31, robodesk/robodesk.py:70-95)."""
    _prepare()
    from gym import spaces
    from mbrl.environments.robodesk_env import Robodesk

    env = object.__new__(Robodesk)
    env.name = "Robodesk"
    env.task, env.reward = task, reward
    env.agent_dim, env.object_dyn_dim, env.object_stat_dim, env.nObj = 24, 13, 6, 8
    env.num_object_types = 6
    O = 24 + 8 * 19
    env.observation_space = spaces.Box(-np.inf, np.inf, shape=(O,), dtype="float32")
    env.action_space = spaces.Box(-1.0, 1.0, shape=(8,), dtype="float32")
    env.observation_space_size_preproc = O
    env.original_pos_z_dict = {"ball": 0.79963282, "upright_block": 0.84978449, "flat_block": 0.77478449}
    env.block_ids = {"ball": 89, "upright_block": 102, "flat_block": 115}
    env.goal = np.zeros(0, np.float32)
    return env


_TRAIN_PARAMS = dict(batch_size=125, epochs=25, iterations=0, learning_rate=1e-5, weight_decay=1e-3,
                     optimizer="Adam", bootstrapped=False, train_epochs_only_with_latest_data=False)


def make_gnn_model(env, hidden_dim=128, num_layers=2, n=5, running_stats=False, seed=0, ignore_agent_object=True,
                   agent_as_global_node=True, embedding_dim=None, num_message_passing=1):
    _prepare()
    from mbrl.models.gnn_ensemble import GNNForwardEnsembleModel

    torch.manual_seed(seed)
    mp = dict(n=n, hidden_dim=hidden_dim, num_layers=num_layers, layer_norm=True, num_message_passing=num_message_passing,
              act_fn="relu", output_act_fn="none", ignore_agent_object=ignore_agent_object)
    if not agent_as_global_node:
        mp.update(embedding_dim=embedding_dim, embedding=True)
    return GNNForwardEnsembleModel(
        env=env,
        model_params=mp,
        train_params=dict(_TRAIN_PARAMS),
        use_input_normalization=True,
        use_output_normalization=True,
        target_is_delta=True,
        normalize_w_running_stats=running_stats,
        agent_as_global_node=agent_as_global_node,
    )


def make_mlp_model(env, hidden_dim=256, num_layers=3, n=5, running_stats=False, seed=0):
    _prepare()
    from mbrl.models.mlp_ensemble import ParallelNNDeterministicEnsemble

    torch.manual_seed(seed)
    return ParallelNNDeterministicEnsemble(
        env=env,
        model_params=dict(n=n, hidden_dim=hidden_dim, num_layers=num_layers, act_fn="silu", output_act_fn="none",
                          layer_norm=False),
        train_params=dict(_TRAIN_PARAMS),
        use_input_normalization=True,
        use_output_normalization=True,
        target_is_delta=True,
        normalize_w_running_stats=running_stats,
    )


def load_into_reference(model, weights, normalizers, running_stats):
    """Seeded weights / normalisers -> a reference model object, through its own load_state_dict and normaliser attributes."""
    model.model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in weights.items()})
    if hasattr(model, "input_normalizer_agent"):
        pairs = {"in_agent": model.input_normalizer_agent, "in_dyn": model.input_normalizer_obj_dyn,
                 "in_stat": model.input_normalizer_obj_stat, "in_action": model.action_normalizer,
                 "out_agent": model.output_agent_normalizer, "out_dyn": model.output_obj_normalizer}
    else:
        pairs = {"in": model.input_normalizer, "out": model.output_normalizer}
    for k, nz in pairs.items():
        m, s = normalizers[k]
        if running_stats:
            nz.mean_tensor, nz.std_tensor = torch.from_numpy(m.copy()), torch.from_numpy(s.copy())
        else:
            nz.mean, nz.std = torch.from_numpy(m.copy()).view(1, -1), torch.from_numpy(s.copy()).view(1, -1)


def make_controller(env, model, *, curiosity, horizon, num_traj, cost_along_trajectory="sum", sampler=None,
                    extrinsic_reward=False, extrinsic_reward_scale=1.0):
    _prepare()
    from mbrl.controllers.mpc_torch import TorchMpcICem
    from mbrl.controllers.mpc_torch_curiosity import TorchCuriosityMpcICem

    sp = dict(opt_iterations=3, elites_size=10, alpha=0.1, init_std=0.8, relative_init=True,
              execute_best_elite=True, keep_previous_elites=True, shift_elites_over_time=True,
              finetune_first_action=False, fraction_elites_reused=0.3, use_mean_actions=True,
              colored_noise=True, noise_beta=3.5, use_ensemble_cost_std=False)
    if sampler:
        sp.update(sampler)
    kw = dict(env=env, forward_model=model, horizon=horizon, num_simulated_trajectories=num_traj,
              factor_decrease_num=1, use_env_reward=False, cost_along_trajectory=cost_along_trajectory,
              action_sampler_params=sp, verbose=False, do_visualize_plan=False, use_async_action=False,
              logging=False)
    if curiosity:
        return TorchCuriosityMpcICem(extrinsic_reward=extrinsic_reward,
                                     extrinsic_reward_scale=extrinsic_reward_scale, **kw)
    return TorchMpcICem(**kw)


class NoiseTap:
    """Replaces colored_noise.torch_powerlaw_psd_gaussian by a replay of pre-drawn noise tensors and records
    the call order -- the RNG boundary of SURVEY.md section 8c."""

    def __init__(self, provider):
        self.provider = provider
        self.calls = []

    def __enter__(self):
        from mbrl.controllers import colored_noise

        self._mod = colored_noise
        self._orig = colored_noise.torch_powerlaw_psd_gaussian

        def fake(exponent, size, fmin=0):
            y = self.provider(exponent, tuple(size))
            self.calls.append((tuple(size), y.clone()))
            return y

        colored_noise.torch_powerlaw_psd_gaussian = fake
        return self

    def __exit__(self, *a):
        self._mod.torch_powerlaw_psd_gaussian = self._orig
