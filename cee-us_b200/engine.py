"""Owner of one libceeus_b200 handle plus the PyTorch-owned device buffers it borrows.

PyTorch is plumbing here (device memory, streams, torch.distributed); all arithmetic happens in the CUDA library.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from .envs import describe_env

SAMPLER_DEFAULTS = dict(opt_iterations=3, elites_size=10, alpha=0.1, init_std=0.8, relative_init=True,
                        execute_best_elite=True, keep_previous_elites=True, shift_elites_over_time=True,
                        finetune_first_action=False, fraction_elites_reused=0.3, use_mean_actions=True,
                        colored_noise=True, noise_beta=1, use_ensemble_cost_std=False)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def weight_key_order(model_kind, num_layers, layer_norm, gnn_variant=0, edge_global=False):
    """state_dict keys in the order ceeus_set_weights expects (include/ceeus_b200.h)."""
    if model_kind != "gnn":
        mlps = (("", num_layers, layer_norm),)
    elif gnn_variant == 0:  # GraphNeuralNetworkEnsemble (gnn_modules.py:78-115)
        mlps = tuple((p, num_layers, layer_norm) for p in ("edge_mlp.", "node_mlp.", "global_mlp.") +
                     (("edge_mlp_global.",) if edge_global else ()))
    else:                   # HomogeneousGraphNeuralNetworkEnsemble (:351-405): embeddings / heads are single linear layers
        mlps = (("edge_mlp.", num_layers, layer_norm), ("node_mlp.", num_layers, layer_norm), ("embed_agent.", 0, False),
                ("embed_obj.", 0, False), ("output_head_agent.", 0, False), ("output_head_objdyn.", 0, False))
    keys = []
    for p, nl, lnorm in mlps:
        for l in range(nl + 1):
            keys += [f"{p}layers.{l}.0.weight", f"{p}layers.{l}.0.bias"]
            if l < nl and lnorm:
                keys += [f"{p}layers.{l}.1.gamma", f"{p}layers.{l}.1.beta"]
    return keys


class Engine:
    def __init__(self, *, env, model_kind, ensemble_size, hidden_dim, num_layers, act_fn, layer_norm, aggr_fn="mean",
                 horizon, num_traj, sampler=None, cost_along_trajectory="sum", curiosity=True, extrinsic_reward=False,
                 extrinsic_reward_scale=1.0, num_envs=1, world_size=1, rank=0, precision="fp32", seed=0,
                 device=None, external_cost=None, cost_in_rollout_tail=False, gnn_variant=0, edge_global=False, num_message_passing=1,
                 embedding_dim=0, tc_row_layout="auto"):
        if not torch.cuda.is_available():
            raise RuntimeError("ceeus_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.env_desc = d = describe_env(env)
        sp = dict(SAMPLER_DEFAULTS)
        sp.update(sampler or {})
        if sp.get("finetune_first_action"):
            raise NotImplementedError("finetune_first_action is dead code upstream (mpc_torch.py:361-381) and unsupported")
        self.sampler = sp
        p_total = num_traj * world_size
        num_elites = min(int(sp["elites_size"]), p_total // 2)  # mpc_torch.py:514-519
        if num_elites < 2:
            num_elites = 2
        num_kept = int(num_elites * float(sp["fraction_elites_reused"]))
        need_env = (not curiosity) or extrinsic_reward
        self.external_cost = bool(need_env and (d["external_cost"] if external_cost is None else external_cost))
        if self.external_cost:
            d["env_kind"] = _lib.ENV_EXTERNAL
        if need_env and not d["task_cost_supported"] and not self.external_cost:
            raise NotImplementedError(f"task cost for env kind={d['kind']} case={d['case']} is not built in and the "
                                      "environment has no cost_fn to call")
        if need_env and d["case"] == "Slide" and d["shaped_reward"]:
            raise AssertionError("Shaped reward cannot be used with Slide!")  # as upstream, fpp_construction_env.py:455
        cfg = _lib.CeeusConfig()
        cfg.abi_version = _lib.ABI_VERSION
        cfg.model_kind = _lib.MODEL_GNN if model_kind == "gnn" else _lib.MODEL_MLP
        cfg.ensemble_size, cfg.n_obj = ensemble_size, d["n_obj"]
        cfg.agent_dim, cfg.dyn_dim, cfg.stat_dim = d["agent_dim"], d["dyn_dim"], d["stat_dim"]
        cfg.action_dim, cfg.goal_dim = d["action_dim"], d["goal_dim"]
        cfg.hidden_dim, cfg.num_layers = hidden_dim, num_layers
        cfg.act_fn, cfg.layer_norm = _lib.ACT[act_fn], int(bool(layer_norm))
        cfg.aggr_mean = 1 if aggr_fn == "mean" else 0
        if precision not in _lib.PREC:
            raise ValueError(f"precision must be one of {sorted(_lib.PREC)} ('tc' = fp16 operands / fp32 accumulate on tcgen05; "
                             f"there is no tf32 / bf16 mode), got {precision!r}")
        cfg.precision = _lib.PREC[precision]
        cfg.num_envs, cfg.num_traj, cfg.horizon = num_envs, num_traj, horizon
        cfg.opt_iterations, cfg.num_elites, cfg.num_kept = int(sp["opt_iterations"]), num_elites, num_kept
        cfg.relative_init = int(bool(sp["relative_init"]))
        cfg.use_mean_actions = int(bool(sp["use_mean_actions"]))
        cfg.keep_previous_elites = int(bool(sp["keep_previous_elites"]))
        cfg.shift_elites_over_time = int(bool(sp["shift_elites_over_time"]))
        cfg.execute_best_elite = int(bool(sp["execute_best_elite"]))
        cfg.use_ensemble_cost_std = int(bool(sp["use_ensemble_cost_std"]))
        cfg.colored_noise = int(bool(sp["colored_noise"]))
        cfg.cost_along_trajectory = _lib.COST[cost_along_trajectory]
        cfg.curiosity, cfg.extrinsic_reward = int(bool(curiosity)), int(bool(extrinsic_reward))
        cfg.alpha, cfg.init_std = float(sp["alpha"]), float(sp["init_std"])
        cfg.noise_beta, cfg.extrinsic_scale = float(sp["noise_beta"]), float(extrinsic_reward_scale)
        cfg.env_kind = d["env_kind"] if need_env else _lib.ENV_NONE
        cfg.cost_case, cfg.sparse, cfg.shaped_reward = d["cost_case"], d["sparse"], d["shaped_reward"]
        cfg.threshold, cfg.buffer_threshold = d["threshold"], d["buffer_threshold"]
        for name in ("coord_weights", "buffer_xyz", "cost_thres", "gripper_init"):
            getattr(cfg, name)[:] = [float(v) for v in d[name]]
        cfg.world_size, cfg.rank, cfg.seed = world_size, rank, seed
        cfg.cost_in_rollout_tail = int(bool(cost_in_rollout_tail))
        cfg.tc_row_layout = {"auto": 0, None: 0, "node": 1, "global": 2}[tc_row_layout]
        cfg.gnn_variant, cfg.edge_global = int(gnn_variant), int(bool(edge_global))
        cfg.num_message_passing, cfg.embedding_dim = int(num_message_passing), int(embedding_dim or 0)
        self.gnn_variant, self.edge_global = int(gnn_variant), bool(edge_global)
        self.cfg = cfg
        self.model_kind, self.num_layers, self.layer_norm = model_kind, num_layers, bool(layer_norm)
        self.handle = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ceeus_create(C.byref(cfg), C.byref(self.handle)))
        self.E, self.H, self.U, self.O = ensemble_size, horizon, d["action_dim"], d["obs_dim"]
        self.P, self.nE, self.k, self.n_kept = num_traj, num_envs, num_elites, num_kept
        self.world_size, self.rank = world_size, rank
        self.rec = self.lib.ceeus_cand_record_len(self.handle)
        f32 = dict(device=self.device, dtype=torch.float32)
        nE, P, H, U, E, O, k = self.nE, self.P, self.H, self.U, self.E, self.O, self.k
        self.mean_ = torch.zeros(nE, H, U, **f32)
        self.std_ = torch.ones(nE, H, U, **f32)
        self.actions_ = torch.zeros(nE * P, H, U, **f32)
        # full trajectories [H+1,E,nE*P,O] are materialised only on request (predict_n_steps, elite_samples, externally
        # evaluated task costs): the planner's own rollouts stay in the library's compact scratch and are reduced to costs in
        # the tail of the rollout kernel
        self._traj = None
        self.comm_ready = False
        self._last_obs_host = None
        self._needs_traj = self.external_cost or bool(os.environ.get("CEEUS_NO_FUSE"))
        self.costs_per_model_ = torch.zeros(nE, P, E, **f32)
        self.costs_ = torch.zeros(nE, P, **f32)
        self.elite_actions_ = torch.zeros(nE, k, H, U, **f32)
        self.elite_costs_ = torch.zeros(nE, k, **f32)
        self.elite_cpm_ = torch.zeros(nE, k, E, **f32)
        self.elite_idx_ = torch.zeros(nE, k, device=self.device, dtype=torch.int32)
        self.cand_ = torch.zeros(nE, k, self.rec, **f32)
        self.action_out_ = torch.zeros(nE, U, **f32)
        self.obs_dev_ = torch.zeros(nE, O, **f32)
        self._bind()
        lo, hi = np.ascontiguousarray(d["action_low"]), np.ascontiguousarray(d["action_high"])
        _lib.check(self.lib.ceeus_set_action_bounds(self.handle, lo.ctypes.data_as(C.c_void_p),
                                                    hi.ctypes.data_as(C.c_void_p)), self.handle)
        self.weights_version = None

    def _bind(self):
        if self._needs_traj and self._traj is None:
            self._traj = torch.zeros(self.H + 1, self.E, self.nE * self.P, self.O, device=self.device, dtype=torch.float32)
        b = _lib.CeeusBuffers()
        for name, t in (("mean", self.mean_), ("std", self.std_), ("actions", self.actions_), ("traj", self._traj),
                        ("costs_per_model", self.costs_per_model_), ("costs", self.costs_),
                        ("elite_actions", self.elite_actions_), ("elite_costs", self.elite_costs_),
                        ("elite_cpm", self.elite_cpm_), ("elite_idx", self.elite_idx_), ("cand", self.cand_),
                        ("action_out", self.action_out_), ("obs", self.obs_dev_)):
            setattr(b, name, t.data_ptr() if t is not None else None)
        self._bufs = b
        _lib.check(self.lib.ceeus_bind_buffers(self.handle, C.byref(b)), self.handle)

    @property
    def traj_(self):
        """The full-trajectory buffer [H+1,E,nE*P,O] (allocated on first use; rollout() without `out` writes here)."""
        if self._traj is None:
            self._needs_traj = True
            self._bind()
        return self._traj

    def materialise(self):
        """Full trajectories of the planner's CURRENT population (actions_) from the current observation(s) obs_dev_ -- what
        the reference keeps in all_observations_ / all_next_observations_ after simulate_trajectories (mpc_torch.py:139-157)."""
        if self._last_obs_host is not None:  # plan_step() keeps the observation in the library's own staging buffer
            self.obs_dev_.copy_(torch.from_numpy(self._last_obs_host))
        return self.rollout(self.obs_dev_, self.actions_)

    def close(self):
        if getattr(self, "handle", None):
            self.lib.ceeus_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- hand-off -------------------------------------------------------------------------------------------
    def set_weights(self, state_dict):
        keys = weight_key_order(self.model_kind, self.num_layers, self.layer_norm, self.gnn_variant, self.edge_global)
        n = self.lib.ceeus_num_weight_tensors(self.handle)
        if n != len(keys):
            raise RuntimeError(f"library expects {n} tensors, state_dict layout gives {len(keys)}")
        tensors = []
        r, c = C.c_int32(), C.c_int32()
        for i, key in enumerate(keys):
            t = torch.as_tensor(state_dict[key]).detach().to(self.device, torch.float32).contiguous()
            _lib.check(self.lib.ceeus_weight_tensor_shape(self.handle, i, C.byref(r), C.byref(c)), self.handle)
            want = (self.E, r.value, c.value) if r.value > 1 or t.dim() == 3 else (self.E, c.value)
            if tuple(t.shape) != want and tuple(t.shape) != (self.E, r.value, c.value):
                raise ValueError(f"{key}: shape {tuple(t.shape)} != expected {want}")
            tensors.append(t)
        arr = (C.c_void_p * n)(*[t.data_ptr() for t in tensors])
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ceeus_set_weights(self.handle, arr, n, _stream()), self.handle)
            torch.cuda.current_stream().synchronize()  # `tensors` may be temporaries

    def set_normalizers(self, groups):
        """groups: list of (mean, std) in the order of include/ceeus_b200.h (GNN: in_agent, in_dyn, in_stat, in_action,
        out_agent, out_dyn; MLP: in, out)."""
        flat = np.concatenate([np.concatenate([np.asarray(m, np.float32).ravel(), np.asarray(s, np.float32).ravel()])
                               for m, s in groups]).astype(np.float32)
        flat = np.ascontiguousarray(flat)
        _lib.check(self.lib.ceeus_set_normalizers(self.handle, flat.ctypes.data_as(C.c_void_p), flat.size), self.handle)

    def set_goal(self, goal):
        if self.env_desc["goal_dim"] == 0:  # environments without goal dims in the observation (RoboDesk)
            return
        g = np.ascontiguousarray(np.broadcast_to(np.asarray(goal, np.float32).reshape(-1, self.env_desc["goal_dim"]),
                                                 (self.nE, self.env_desc["goal_dim"])))
        _lib.check(self.lib.ceeus_set_goal(self.handle, g.ctypes.data_as(C.c_void_p)), self.handle)

    # ---- stand-alone stages ------------------------------------------------------------------------------------
    def rollout(self, start_obs, actions, out=None):
        """start_obs [n_start,O] cuda, actions [n,H,U] cuda -> traj [H+1,E,n,O]."""
        n = actions.shape[0]
        start_obs = start_obs.to(self.device, torch.float32).contiguous()
        actions = actions.to(self.device, torch.float32).contiguous()
        if out is None:  # like the reference's preallocated all_observations_ buffers: reused across calls
            out = self.traj_ if n == self.nE * self.P else torch.empty(self.H + 1, self.E, n, self.O,
                                                                          device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ceeus_rollout(self.handle, _ptr(start_obs), start_obs.shape[0], _ptr(actions), n,
                                              _ptr(out), _stream()), self.handle)
        return out

    def costs(self, traj):
        n = traj.shape[2]
        cpm = torch.empty(n, self.E, device=self.device, dtype=torch.float32)
        costs = torch.empty(n, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ceeus_costs(self.handle, _ptr(traj), n, _ptr(cpm), _ptr(costs), _stream()), self.handle)
        return cpm, costs

    def sample(self, mean, std, n, noise=None, gauss=None, write_gauss=False, stream_id=0, first_row=0):
        actions = torch.empty(n, self.H, self.U, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ceeus_sample(self.handle, _ptr(mean), _ptr(std), _ptr(noise), _ptr(gauss),
                                             int(write_gauss), stream_id, first_row, n, _ptr(actions), _stream()),
                       self.handle)
        return actions

    # ---- planner --------------------------------------------------------------------------------------------------
    def begin_rollout(self):
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ceeus_begin_rollout(self.handle, _stream()), self.handle)

    def has_elites(self):
        return bool(self.lib.ceeus_has_elites(self.handle))

    def plan_step(self, obs, noise=None):
        """obs: host float32 [nE,O] (or [O]); returns host float32 [nE,U].  One H2D + one D2H inside."""
        obs = np.ascontiguousarray(np.asarray(obs, np.float32).reshape(self.nE, self.O))
        self._last_obs_host = obs
        out = np.empty((self.nE, self.U), np.float32)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ceeus_plan_step(self.handle, obs.ctypes.data_as(C.c_void_p),
                                                out.ctypes.data_as(C.c_void_p), _ptr(noise), _stream()), self.handle)
        return out

    def plan_step_resident(self, noise=None):
        """One MPC step with the observation already in ``obs_dev_`` and the action left in ``action_out_``: one CUDA-graph
        launch (device RNG), no host copy, no host synchronisation."""
        self._last_obs_host = None
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ceeus_plan_step_resident(self.handle, _ptr(noise), _stream()), self.handle)
        return self.action_out_

    def comm_init(self, group=None, p2p=True):
        """Collective over the ranks of ``group``.  On one node (<= 8 ranks) the ranks map each other's candidate buffers through
        CUDA IPC (the handles travel through torch.distributed) and exchange by direct NVLink stores; otherwise / on failure the
        library creates an NCCL communicator.  Afterwards plan_step / plan_step_resident are one graph launch at any world size."""
        import torch.distributed as dist

        if p2p and 2 <= self.world_size <= 8:
            buf = (C.c_uint8 * 64)()
            with torch.cuda.device(self.device):
                rc = self.lib.ceeus_p2p_export(self.handle, buf)
            mine = torch.tensor(list(buf), dtype=torch.uint8)
            ok = torch.tensor([1 if rc == 0 else 0], dtype=torch.int32)
            on_dev = dist.get_backend(group) == "nccl"
            allh = torch.zeros(self.world_size * 64, dtype=torch.uint8, device=self.device if on_dev else "cpu")
            dist.all_gather_into_tensor(allh, mine.to(allh.device), group=group)
            raw = (C.c_uint8 * (64 * self.world_size))(*allh.cpu().tolist())
            if rc == 0:
                with torch.cuda.device(self.device):
                    rc = self.lib.ceeus_p2p_import(self.handle, raw)
            ok[0] = 1 if rc == 0 else 0
            ok = ok.to(allh.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)  # all ranks take the same path
            if int(ok.item()) == 1:
                self.comm_ready = True
                return

        ident = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            buf = (C.c_uint8 * 128)()
            _lib.check(self.lib.ceeus_comm_unique_id(buf), None)
            ident = torch.tensor(list(buf), dtype=torch.uint8)
        ident = ident.to(self.device) if dist.get_backend(group) == "nccl" else ident
        dist.broadcast(ident, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        raw = (C.c_uint8 * 128)(*ident.cpu().tolist())
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ceeus_comm_init(self.handle, raw), self.handle)
        self.comm_ready = True

    def plan_iter_local(self, it, noise=None, shift_noise=None):
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ceeus_plan_iter_local(self.handle, it, _ptr(self.obs_dev_), _ptr(noise),
                                                      _ptr(shift_noise), _stream()), self.handle)

    def plan_iter_rollout(self, it, noise=None, shift_noise=None):
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ceeus_plan_iter_rollout(self.handle, it, _ptr(self.obs_dev_), _ptr(noise),
                                                        _ptr(shift_noise), _stream()), self.handle)

    def plan_iter_costs(self, env_cost=None):
        """env_cost: [nE*P,E,H] float32 cuda tensor (env.cost_fn on the materialised trajectories) or None."""
        if env_cost is not None:
            env_cost = env_cost.to(self.device, torch.float32).contiguous()
            if tuple(env_cost.shape) != (self.nE * self.P, self.E, self.H):
                raise ValueError(f"env.cost_fn returned {tuple(env_cost.shape)}, expected {(self.nE * self.P, self.E, self.H)}")
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ceeus_plan_iter_costs(self.handle, _ptr(env_cost), _stream()), self.handle)

    def plan_iter_update(self, it, cand_all):
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ceeus_plan_iter_update(self.handle, it, _ptr(cand_all), _stream()), self.handle)

    def plan_finish(self):
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ceeus_plan_finish(self.handle, _stream()), self.handle)

    def tc_status(self):
        return int(self.lib.ceeus_tc_status(self.handle))

    def profile_rollouts(self, enable):
        """Switches the event pairs around the rollout launches on / off; returns the durations (ms) recorded so far."""
        buf = (C.c_float * 256)()
        n = self.lib.ceeus_profile_rollouts(self.handle, int(bool(enable)), buf, 256)
        if n < 0:
            _lib.check(n, self.handle)
        return [float(buf[i]) for i in range(n)]

    def tc_tiling(self, n):
        """(rounds[E], uneven[E]) of a tensor-core launch over n trajectories (ceeus_tc_tiling)."""
        r, u = (C.c_int32 * self.E)(), (C.c_int32 * self.E)()
        _lib.check(self.lib.ceeus_tc_tiling(self.handle, int(n), r, u), self.handle)
        return list(r), list(u)

    def tc_row_layout(self):
        """'node' / 'global': the tile row layout of the tensor-core GNN rollout of this handle (ceeus_tc_row_layout)."""
        return {0: None, 1: "node", 2: "global"}[int(self.lib.ceeus_tc_row_layout(self.handle))]

    def kernel_launches(self):
        return int(self.lib.ceeus_kernel_launches(self.handle))
