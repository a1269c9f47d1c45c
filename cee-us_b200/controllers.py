"""Drop-in iCEM controllers: ``TorchMpcICem`` (yaml ``controller: mpc-icem-torch``, mbrl/controllers/mpc_torch.py:20-526)
and ``TorchCuriosityMpcICem`` (``mpc-curiosity-icem-torch``, mbrl/controllers/mpc_torch_curiosity.py:9-95) with the
same constructor keywords (yaml ``controller_params``), the same ``beginning_of_rollout / get_action / end_of_rollout``
surface and the same readable attributes (``mean_, std_, costs_, elite_samples, horizon, num_sim_traj``).

``get_action`` is ONE call into libceeus_b200 (ceeus_plan_step): obs goes host->device, opt_iterations x
{sample, rollout, costs, top-k, refit} run as CUDA kernels without returning to Python, the action comes back.
With ``torch.distributed`` initialised (one process per GPU) the trajectories are sharded over ranks and each
iteration adds one small all-gather (distributed.py).
"""
import numpy as np
import torch
import torch.distributed as dist

from .distributed import gather_candidates
from .engine import Engine
from .rollout_buffer import RolloutView


class B200MpcICem:
    """yaml key ``mpc-icem-b200`` (task cost = env.cost_fn evaluated on device)."""

    _curiosity = False

    def __init__(self, *, env, forward_model, horizon, num_simulated_trajectories, action_sampler_params,
                 cost_along_trajectory="sum", use_env_reward=False, factor_decrease_num=1, hook=None, hook_params=None,
                 verbose=False, do_visualize_plan=False, use_async_action=False, logging=True,
                 fully_deterministic=False, backend_params=None, **kwargs):
        if use_env_reward:
            raise NotImplementedError()  # as upstream: mpc_torch.py:118-119
        if use_async_action:
            raise NotImplementedError("use_async_action is a real-robot feature (rollout_utils.py:257-262)")
        if hook:
            raise NotImplementedError("mpc hooks (cem_memory.py) are out of scope; `hook: null` in all shipped settings")
        if num_simulated_trajectories < 2:
            raise ValueError("At least two trajectories needed!")  # mpc.py:41-42
        if kwargs:
            raise TypeError(f"unexpected controller_params: {sorted(kwargs)}")
        self.env, self.forward_model = env, forward_model
        self.horizon, self.num_sim_traj = horizon, num_simulated_trajectories
        self.factor_decrease_num = factor_decrease_num
        self.cost_along_trajectory = cost_along_trajectory
        self.action_sampler_params = dict(action_sampler_params)
        self.verbose, self.logging, self.do_visualize_plan = verbose, logging, do_visualize_plan
        self.fully_deterministic = fully_deterministic  # the device top-k is always the stable order (cost, index)
        self.backend_params = dict(backend_params or {})
        self._extra_engine_kwargs = {}
        self.was_reset = False
        self.forward_model_state = None
        self.mpc_hook = None
        self._ensemble_size = forward_model.ensemble_size
        self.forward_model.num_simulated_trajectories = self.num_sim_traj
        self.forward_model.horizon = self.horizon
        self.forward_model.preallocate_memory()
        self._engine = None
        self._last_obs = None
        self._iterated = False
        self.preallocate_memory()

    # -- engine -------------------------------------------------------------------------------------------------
    def _world(self):
        if self.backend_params.get("distributed", True) and dist.is_available() and dist.is_initialized():
            return dist.get_world_size(), dist.get_rank()
        return 1, 0

    def preallocate_memory(self):
        """All planner buffers are allocated once here and bound to the library (mpc_torch.py:64-110)."""
        if self._engine is not None:
            self._engine.close()
        world, rank = self._world()
        if self.num_sim_traj % world != 0:
            raise ValueError("num_simulated_trajectories must be divisible by the number of GPUs")
        kw = self.forward_model.engine_kwargs()
        kw["precision"] = self.backend_params.get("precision", kw.get("precision", "fp32"))
        kw["tc_row_layout"] = self.backend_params.get("tc_row_layout", kw.get("tc_row_layout", "auto"))
        self._engine = Engine(horizon=self.horizon, num_traj=self.num_sim_traj // world,
                              sampler=self.action_sampler_params, cost_along_trajectory=self.cost_along_trajectory,
                              curiosity=self._curiosity, num_envs=self.backend_params.get("num_envs", 1),
                              world_size=world, rank=rank, seed=self.backend_params.get("seed", 0),
                              cost_in_rollout_tail=self.backend_params.get("cost_in_rollout_tail", False),
                              **self._extra_engine_kwargs, **kw)
        e = self._engine
        if world > 1 and self.backend_params.get("c_comm", True) and not e.external_cost:
            e.comm_init()  # library-side NCCL communicator: the N-rank step is one CUDA-graph launch, like the 1-GPU step
        self.num_elites = e.k
        self.opt_iter = int(e.sampler["opt_iterations"])
        self.mean_, self.std_ = e.mean_[0], e.std_[0]          # [H,U] views of the bound buffers
        self.costs_ = e.costs_[0]                              # [P_local] costs of this rank's fresh samples
        self.costs_per_model_ = e.costs_per_model_[0]

    @property
    def engine(self):
        return self._engine

    @property
    def elite_samples(self):
        """Elites of the last iteration as a SimpleRolloutBuffer-like object; the predicted observations are
        re-simulated on demand (nothing on the planning path reads them)."""
        e = self._engine
        if not self.was_reset or not (e.has_elites() or self._iterated):
            return None
        acts = e.elite_actions_[0]
        model = self.forward_model
        last_obs = self._last_obs  # the observation the elites were planned from (the library keeps its own device copy)

        class _Lazy(RolloutView):
            def _materialise(inner):
                if "observations" not in inner.buffer:
                    eng = model._rollout_engine(acts.shape[0], e.H)
                    start = torch.from_numpy(np.ascontiguousarray(last_obs.reshape(-1, e.O)[:1])).to(e.device)
                    traj = eng.rollout(start, acts.contiguous())
                    full = RolloutView.from_device_buffers(traj.clone(), acts)
                    inner.buffer.update(full.buffer)

            def as_array(inner, key):
                if key != "actions":
                    inner._materialise()
                return inner.buffer[key]

            def __getitem__(inner, item):
                if not (isinstance(item, str) and item == "actions"):
                    inner._materialise()
                return RolloutView.__getitem__(inner, item)

        return _Lazy(actions=acts.unsqueeze(1).expand(-1, e.E, -1, -1))

    # -- Controller surface (base_types.py:40-58, abstract_controller.py:47-62) --------------------------------
    @torch.no_grad()
    def beginning_of_rollout(self, *, observation, state=None, mode="train"):
        self.forward_model_state = self.forward_model.reset(observation)
        self._engine.begin_rollout()
        self.was_reset = True
        self._iterated = False
        self._last_obs = np.asarray(observation, np.float32).copy()

    def end_of_rollout(self, total_time=None, total_return=None, mode="train"):
        pass

    @torch.no_grad()
    def get_action(self, obs, state=None, mode="train", noise=None):
        """obs: np.ndarray [O] (or [num_envs,O]) -> np.ndarray [U] (or [num_envs,U]).

        ``noise`` (testing / parity hook): device tensor holding this rank's colored-noise draws in the reference's
        order (see ceeus_plan_step in include/ceeus_b200.h); ``None`` = on-device Philox sampler."""
        if not self.was_reset:
            raise AttributeError("beginning_of_rollout() needs to be called before")
        e = self._engine
        self.forward_model_state = self.forward_model.got_actual_observation_and_env_state(
            observation=obs, env_state=state, model_state=self.forward_model_state)
        self.forward_model.push_to(e)
        e.set_goal(np.asarray(self.env.goal, np.float32))  # env.goal changes on every env.reset (obs_postproc)
        obs = np.asarray(obs, np.float32)
        single = obs.ndim == 1
        self._last_obs = obs.copy()
        if (e.world_size == 1 or e.comm_ready) and not e.external_cost:
            action = e.plan_step(obs, noise)
        else:
            action = self._distributed_step(obs, noise)
        self._iterated = True
        if self.logging and self.backend_params.get("log_fn"):
            self.backend_params["log_fn"](float(e.elite_costs_[0, 0]), "best_trajectory_cost")
        return action[0] if single else action

    @torch.no_grad()
    def plan_on_device(self, noise=None):
        """The same MPC step with the observation already resident in ``engine.obs_dev_`` and the action left in
        ``engine.action_out_``: no host<->device copy, no host synchronisation (used by bench.py's device-timed leg
        and by callers that keep their environments on the GPU)."""
        e = self._engine
        self.forward_model.push_to(e)
        if (e.world_size == 1 or e.comm_ready) and not e.external_cost:
            self._iterated = True
            return e.plan_step_resident(noise)
        per_iter = e.nE * e.P * e.U * e.H
        per_shift = e.nE * e.n_kept * e.U * e.H
        off = 0
        shifting = bool(e.cfg.shift_elites_over_time) and e.has_elites() and e.n_kept > 0
        for i in range(self.opt_iter):
            ni = ns = None
            if noise is not None:
                flat = noise.reshape(-1)
                ni = flat[off:off + per_iter]
                off += per_iter
                if i == 0 and shifting:
                    ns = flat[off:off + per_shift]
                    off += per_shift
            if e.external_cost:
                # the environment's own cost_fn on the materialised rollout (mpc_torch.py:121-125); everything else stays on
                # the device path
                e.plan_iter_rollout(i, ni, ns)
                e.plan_iter_costs(self._cost_from_rollout(RolloutView.from_device_buffers(e.traj_, e.actions_)))
            else:
                e.plan_iter_local(i, ni, ns)
            e.plan_iter_update(i, gather_candidates(e.cand_).contiguous() if e.world_size > 1 else e.cand_)
        e.plan_finish()
        self._iterated = True
        return e.action_out_

    def _cost_from_rollout(self, view):
        """Costs [P,E,H] of a materialised rollout for environments whose task cost the library does not know: the environment's
        own cost_fn (mpc_torch.py:121-125); subclasses put other caller-side costs here (rnd.py)."""
        return self.env.cost_fn(view["observations"], view["actions"], view["next_observations"])

    def _distributed_step(self, obs, noise):
        e = self._engine
        e.obs_dev_.copy_(torch.from_numpy(np.ascontiguousarray(obs.reshape(e.nE, e.O))), non_blocking=True)
        e._last_obs_host = None
        return self.plan_on_device(noise).cpu().numpy()

    def reset_horizon(self, horizon):  # mpc_torch_curiosity.py:82-95
        if horizon == self.horizon:
            return
        self.horizon = horizon
        self.forward_model.horizon = horizon
        self.preallocate_memory()
        self.was_reset = False


class B200CuriosityMpcICem(B200MpcICem):
    """yaml key ``mpc-curiosity-icem-b200``: cost = -ensemble disagreement (+ optional extrinsic task cost)."""

    _curiosity = True

    def __init__(self, *, object_centric=False, extrinsic_reward=False, extrinsic_reward_scale=1.0, **kwargs):
        self._w_object_centric = object_centric
        self._w_extrinsic_reward = extrinsic_reward
        self._maybe_extrinsic_reward_scale = extrinsic_reward_scale
        self._pending = dict(extrinsic_reward=extrinsic_reward, extrinsic_reward_scale=extrinsic_reward_scale)
        super().__init__(**kwargs)

    def preallocate_memory(self):
        self._extra_engine_kwargs = dict(self._pending)
        super().preallocate_memory()
