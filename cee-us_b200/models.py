"""Drop-in forward models: the inference half of

  * ``GNNForwardEnsembleModel``          (yaml ``forward_model: ParallelGNNDeterministicEnsemble``,
                                          mbrl/models/gnn_ensemble.py:32-983)
  * ``ParallelNNDeterministicEnsemble``  (mbrl/models/mlp_ensemble.py:24-641)

behind the same constructor keywords (the yaml ``forward_model_params``), the same ``predict_n_steps`` /
``predict`` / ``reset`` / ``load`` / ``state_dict`` surface and the same checkpoint format; the rollout itself is ONE
launch of the CUDA kernel in cee-us_b200/csrc.  Training stays with the reference (out of scope):
after ``train()`` hand the new weights over with ``load_state_dict`` / ``sync_from``.
"""
import numpy as np
import torch

from .engine import Engine
from .rollout_buffer import RolloutView


def _np(x):
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    return np.asarray(x)


class _B200Ensemble:
    model_kind = None
    _norm_groups = ()

    def __init__(self, *, env, model_params, train_params=None, target_is_delta=True, use_input_normalization=True,
                 use_output_normalization=True, normalize_w_running_stats=True, agent_as_global_node=True,
                 backend_params=None, **kwargs):
        if not target_is_delta:
            raise NotImplementedError("target_is_delta=False is not used by any shipped setting and is unsupported")
        self.env = env
        # agent_as_global_node: false (or no agent features) -> HomogeneousGraphNeuralNetworkEnsemble (gnn_ensemble.py:62-66, 116-132)
        self.agent_as_global_node = bool(agent_as_global_node)
        self.train_params = dict(train_params or {})
        self.target_is_delta = target_is_delta
        self.use_input_normalization = use_input_normalization
        self.use_output_normalization = use_output_normalization
        self.normalize_w_running_stats = normalize_w_running_stats
        self.backend_params = dict(backend_params or {})
        self._parse_model_params(**model_params)
        self.num_simulated_trajectories = None
        self.horizon = None
        self.memory_is_preallocated = False
        self._state = None          # state_dict (reference key names) of torch tensors
        self._normalizers = None    # name -> (mean, std) numpy
        self._norm_extra = {}       # name -> remaining fields of a loaded normaliser state_dict (running sums, count)
        self._ckpt_extra = {}       # other checkpoint entries carried through load() -> save() unchanged (optimizer)
        self.weights_version = 0
        self._engines = {}
        self._reset_normalizers()

    # -- parameters ---------------------------------------------------------------------------------------
    def _reset_normalizers(self):
        self._normalizers = {k: (np.zeros(d, np.float32), np.ones(d, np.float32)) for k, d in self._norm_dims().items()}

    def load_state_dict(self, state_dict):
        self._state = {k: torch.as_tensor(v).detach().clone().float() for k, v in state_dict.items()}
        self.weights_version += 1

    def state_dict(self):
        if self._state is None:
            raise RuntimeError("no weights loaded")
        return self._state

    def set_normalizers(self, normalizers):
        """normalizers: name -> (mean, std); names as in ``_norm_groups``.  Disabled normalisation == identity."""
        for k, (m, s) in normalizers.items():
            if k not in self._normalizers:
                raise KeyError(k)
            m, s = _np(m).astype(np.float32).reshape(-1), _np(s).astype(np.float32).reshape(-1)
            want = self._normalizers[k][0].shape[0]
            if m.size == 1 and want != 1:  # an un-updated running Normalizer holds scalars (torch_helpers.py:812-816)
                m, s = np.full(want, m[0], np.float32), np.full(want, s[0], np.float32)
            assert m.shape[0] == want and s.shape[0] == want, (k, m.shape, want)
            self._normalizers[k] = (m, s)
        self.weights_version += 1

    # -- checkpoint format of the reference (gnn_ensemble.py:756-808, mlp_ensemble.py:613-638) -------------------
    def _normalizer_state(self, k):
        """state_dict of normaliser group ``k`` in the reference's format: Normalizer_v2 {mean, std} as [1,dim] arrays
        (torch_helpers.py:793-797); running Normalizer {mean, std, sum, sum_of_squares, count} (:870-877) -- the running
        sums read by load() are written back unchanged so that a reference model continues its statistics."""
        m, s = self._normalizers[k]
        if self.normalize_w_running_stats:
            out = {"mean": m.astype(np.float64), "std": s.astype(np.float64)}
            out.update(self._norm_extra.get(k, {}))
            return out
        return {"mean": m.reshape(1, -1).copy(), "std": s.reshape(1, -1).copy()}

    def _read_normalizer_state(self, k, sd):
        self._norm_extra[k] = {f: v for f, v in sd.items() if f not in ("mean", "std")}
        return sd["mean"], sd["std"]

    def _normalizer_list(self):
        out = []
        for k in self._norm_groups:
            m, s = self._normalizers[k]
            is_input = k.startswith("in")
            if (is_input and not self.use_input_normalization) or (not is_input and not self.use_output_normalization):
                m, s = np.zeros_like(m), np.ones_like(s)
            out.append((m, s))
        return out

    def sync_from(self, reference_model):
        """Weight hand-off from a live reference model object (after its ``train()`` / ``load()``)."""
        self.load_state_dict(reference_model.model.state_dict())
        self.set_normalizers(self._read_reference_normalizers(reference_model))

    @staticmethod
    def _nz(n):
        if hasattr(n, "mean_tensor"):  # Normalizer (running stats), torch_helpers.py:804-883
            return _np(n.mean_tensor), _np(n.std_tensor)
        return _np(n.mean), _np(n.std)  # Normalizer_v2, :767-801

    # -- planner-facing surface (base_types.py:61-118) -----------------------------------------------------
    def preallocate_memory(self):
        self.memory_is_preallocated = True

    def got_actual_observation_and_env_state(self, *, observation, env_state=None, model_state=None):
        return None

    def engine_kwargs(self):
        return dict(env=self.env, model_kind=self.model_kind, ensemble_size=self.ensemble_size,
                    hidden_dim=self.hidden_dim, num_layers=self.num_layers, act_fn=self.act_fn,
                    layer_norm=self.layer_norm, aggr_fn=getattr(self, "aggr_fn", "mean"),
                    precision=self.backend_params.get("precision", "fp32"),
                    tc_row_layout=self.backend_params.get("tc_row_layout", "auto"), **self._variant_kwargs())

    def _variant_kwargs(self):
        return {}

    def push_to(self, engine):
        """(Re)uploads weights, normalisers if ``engine`` holds an older version."""
        if engine.weights_version != self.weights_version:
            engine.set_weights(self.state_dict())
            engine.set_normalizers(self._normalizer_list())
            engine.weights_version = self.weights_version

    def _rollout_engine(self, n, horizon):
        key = (n, horizon)
        if key not in self._engines:
            if len(self._engines) > 4:
                self._engines.pop(next(iter(self._engines))).close()
            self._engines[key] = Engine(horizon=horizon, num_traj=max(n, 2), **self.engine_kwargs())
        eng = self._engines[key]
        self.push_to(eng)
        eng.set_goal(np.asarray(self.env.goal, np.float32))
        return eng

    @torch.no_grad()
    def predict_n_steps(self, *, start_observations, start_states=None, policy, horizon):
        """-> (buffer with observations / next_observations / actions [P,E,H,*], None)   gnn_ensemble.py:946-983"""
        acts = getattr(policy, "action_sequences", None)
        if acts is None:
            raise NotImplementedError("only open-loop policies (OpenLoopPolicy.action_sequences [P,H,U]) are supported")
        acts = torch.as_tensor(acts)
        if start_observations.ndim != 2:
            raise AttributeError("call predict_n_steps with a batch of observations [P, obs_dim]")
        eng = self._rollout_engine(acts.shape[0], horizon)
        start = torch.as_tensor(start_observations).to(eng.device, torch.float32)
        acts = acts.to(eng.device, torch.float32).contiguous()
        traj = eng.rollout(start.contiguous(), acts)
        return RolloutView.from_device_buffers(traj, acts), None

    @torch.no_grad()
    def predict(self, *, observations, states=None, actions):
        """One model step for every member: observations [E,B,O] (or [O]), actions [E,B,U] -> (next_obs, None, None)."""
        obs = torch.as_tensor(observations, dtype=torch.float32)
        act = torch.as_tensor(actions, dtype=torch.float32)
        if obs.ndim == 1:
            obs, act = obs[None, None].expand(self.ensemble_size, 1, -1), act[None, None].expand(self.ensemble_size, 1, -1)
        E, B, O = obs.shape
        # member e must start from its own state: run E*B single-step rollouts and take the diagonal blocks
        eng = self._rollout_engine(E * B, 1)
        start = obs.reshape(E * B, O).to(eng.device).contiguous()
        a = act.reshape(E * B, 1, -1).to(eng.device).contiguous()
        traj = eng.rollout(start, a)  # [2,E,E*B,O]
        nxt = traj[1].view(E, E, B, O)
        idx = torch.arange(E, device=nxt.device)
        return nxt[idx, idx], None, None

    def rollout_field_names(self):
        return "observations", "next_observations", "actions", "rewards"  # abstract_models.py:29-30

    @torch.no_grad()
    def rollout_generator(self, start_states, start_observations, horizon, policy, mode=None):
        """Like the reference's (gnn_ensemble.py:876-901, mlp_ensemble.py:502-528; not a generator there either): fills
        ``all_observations_ / all_next_observations_ / all_actions_`` ([H,E,P,*]) -- here as views of one CUDA rollout."""
        buf, _ = self.predict_n_steps(start_observations=torch.as_tensor(start_observations), start_states=start_states,
                                      policy=policy, horizon=horizon)
        self.all_observations_ = buf.as_array("observations").permute(2, 1, 0, 3)
        self.all_next_observations_ = buf.as_array("next_observations").permute(2, 1, 0, 3)
        self.all_actions_ = buf.as_array("actions").permute(2, 1, 0, 3)

    def get_state(self):
        return None  # base_types.py:105-109

    def set_state(self, state):
        pass

    # -- training (training.py) -----------------------------------------------------------------
    default_bootstrapped = True   # gnn_ensemble.py:732 (the dense model defaults to False, mlp_ensemble.py:124)

    @property
    def epsilon(self):
        return float(self.train_params.get("epsilon", 1e-6))

    def _normalizer_list_by_name(self):
        return dict(zip(self._norm_groups, self._normalizer_list()))

    def train(self, rollout_buffer, eval_buffer=None, maybe_update_normalizer=True):
        """Batched training of all members on the CUDA device with the reference's semantics (normaliser update, per-member
        batches, MSE summed over members, Adam eps=1e-4); the planner's engines pick the new weights up on their next call."""
        from .training import EnsembleTrainer

        if self._state is None:
            raise RuntimeError("no weights to train: load_state_dict / load / sync_from first")
        if getattr(self, "_trainer", None) is None or self._trainer_version != self.weights_version:
            self._trainer = EnsembleTrainer(self)   # (re)built whenever the weights were replaced from outside
        stats = self._trainer.train(rollout_buffer, eval_buffer, maybe_update_normalizer)
        self._trainer_version = self.weights_version
        return stats


class B200GNNEnsemble(_B200Ensemble):
    """yaml key ``ParallelGNNDeterministicEnsembleB200``."""

    model_kind = "gnn"
    _norm_groups = ("in_agent", "in_dyn", "in_stat", "in_action", "out_agent", "out_dyn")

    def _parse_model_params(self, *, n, hidden_dim, act_fn="relu", output_act_fn="none", num_layers=1,
                            ignore_agent_object=True, num_message_passing=1, aggr_fn="mean", layer_norm=True,
                            embedding_dim=None, embedding=False):
        if output_act_fn not in ("none", None):
            raise NotImplementedError("output_act_fn other than 'none'")
        self.ensemble_size, self.hidden_dim, self.act_fn, self.num_layers = n, hidden_dim, act_fn, num_layers
        self.layer_norm, self.aggr_fn = layer_norm, aggr_fn
        self.ignore_agent_object, self.num_message_passing = bool(ignore_agent_object), int(num_message_passing)
        if self.agent_as_global_node:
            # like the reference: embedding settings are dropped, and num_message_passing is NOT handed to
            # GraphNeuralNetworkEnsemble (gnn_ensemble.py:100-114), i.e. the global-node variant always does one round
            self.embedding_dim, self.embedding = None, False
        else:
            if self.env.agent_dim <= 0:
                raise NotImplementedError("homogeneous GNN without agent features (embedding optional) is not built")
            if not embedding_dim:
                raise ValueError("agent_as_global_node=False needs model_params.embedding_dim")
            self.embedding_dim, self.embedding = int(embedding_dim), True  # forced on whenever agent_dim > 0 (gnn_modules.py:322-323)

    def _variant_kwargs(self):
        if self.agent_as_global_node:
            return dict(gnn_variant=0, edge_global=not self.ignore_agent_object)
        return dict(gnn_variant=1, num_message_passing=self.num_message_passing, embedding_dim=self.embedding_dim)

    def _norm_dims(self):
        e = self.env
        return {"in_agent": e.agent_dim, "in_dyn": e.object_dyn_dim, "in_stat": e.object_stat_dim,
                "in_action": e.action_space.shape[0], "out_agent": e.agent_dim, "out_dyn": e.object_dyn_dim}

    def _read_reference_normalizers(self, m):
        return {"in_agent": self._nz(m.input_normalizer_agent), "in_dyn": self._nz(m.input_normalizer_obj_dyn),
                "in_stat": self._nz(m.input_normalizer_obj_stat), "in_action": self._nz(m.action_normalizer),
                "out_agent": self._nz(m.output_agent_normalizer), "out_dyn": self._nz(m.output_obj_normalizer)}

    _CKPT = {"in_agent": "input_normalizer_agent", "in_dyn": "input_normalizer_obj_dyn",
             "in_stat": "input_normalizer_obj_stat", "in_action": "action_normalizer",
             "out_agent": "output_agent_normalizer", "out_dyn": "output_obj_normalizer"}

    def reset(self, obs):
        return {"obs": torch.as_tensor(np.asarray(obs), dtype=torch.float32)}  # gnn_ensemble.py:810-830

    def update_normalizer(self, batch):
        """GNNForwardEnsembleModel.update_normalizer (gnn_ensemble.py:428-487): statistics of the agent / object / action
        inputs and of the delta targets over the given transitions."""
        from .training import update_normalizer_state

        if not (self.use_input_normalization or self.use_output_normalization):
            return
        obs, act, nxt = (np.asarray(batch[k], np.float32) for k in ("observations", "actions", "next_observations"))
        e = self.env
        A, D, S, N = e.agent_dim, e.object_dyn_dim, e.object_stat_dim, e.nObj
        dyn = lambda o: o[:, A:A + N * D].reshape(-1, D)
        if self.use_input_normalization:
            update_normalizer_state(self, "in_agent", obs[:, :A])
            update_normalizer_state(self, "in_dyn", dyn(obs))
            update_normalizer_state(self, "in_stat", obs[:, A + N * D:A + N * (D + S)].reshape(-1, S))
            update_normalizer_state(self, "in_action", act)
        if self.use_output_normalization:
            update_normalizer_state(self, "out_agent", nxt[:, :A] - obs[:, :A])
            update_normalizer_state(self, "out_dyn", dyn(nxt) - dyn(obs))

    def load(self, path, for_training=True):
        """Reads a checkpoint written by GNNForwardEnsembleModel.save (gnn_ensemble.py:756-808)."""
        sd = torch.load(path, map_location="cpu", weights_only=False)
        self.load_state_dict(sd["model"])
        self.set_normalizers({k: self._read_normalizer_state(k, sd[v]) for k, v in self._CKPT.items() if v in sd})
        self._ckpt_extra = {k: v for k, v in sd.items() if k == "optimizer"}

    def save(self, path):
        """Writes the reference's checkpoint format: normaliser entries only for the enabled groups (:758-786)."""
        out = {"model": {k: v.cpu() for k, v in self.state_dict().items()}}
        for k, v in self._CKPT.items():
            if self.use_input_normalization if k.startswith("in") else self.use_output_normalization:
                out[v] = self._normalizer_state(k)
        out.update(self._ckpt_extra)
        torch.save(out, path)


class B200MLPEnsemble(_B200Ensemble):
    """yaml key ``ParallelNNDeterministicEnsembleB200``."""

    model_kind = "mlp"
    _norm_groups = ("in", "out")

    def _parse_model_params(self, *, n, hidden_dim, act_fn="silu", output_act_fn="none", num_layers=1, layer_norm=False,
                            spectral_norm=False, weight_initializer="torch_truncated_normal",
                            bias_initializer="constant_zero", l1_reg=0, l2_reg=0):
        if output_act_fn not in ("none", None) or spectral_norm:
            raise NotImplementedError("output_act_fn / spectral_norm")
        self.ensemble_size, self.hidden_dim, self.act_fn, self.num_layers = n, hidden_dim, act_fn, num_layers
        self.layer_norm = layer_norm

    def _norm_dims(self):
        sd = self.env.observation_space_size_preproc
        return {"in": sd + self.env.action_space.shape[0], "out": sd}

    def _read_reference_normalizers(self, m):
        return {"in": self._nz(m.input_normalizer), "out": self._nz(m.output_normalizer)}

    default_bootstrapped = False

    def reset(self, observation):
        return None  # mlp_ensemble.py:640-641

    def update_normalizer(self, batch):
        """ParallelNNDeterministicEnsemble.update_normalizer (mlp_ensemble.py:172-188)."""
        from .training import update_normalizer_state

        obs, act, nxt = (np.asarray(batch[k], np.float32) for k in ("observations", "actions", "next_observations"))
        sd = self.env.observation_space_size_preproc
        if self.use_input_normalization:
            update_normalizer_state(self, "in", np.concatenate([obs[:, :sd], act], -1))
        if self.use_output_normalization:
            update_normalizer_state(self, "out", nxt[:, :sd] - obs[:, :sd])

    def load(self, path, for_training=True):
        sd = torch.load(path, map_location="cpu", weights_only=False)  # mlp_ensemble.py:625-638
        self.load_state_dict(sd["nn_state"])
        self.set_normalizers({"in": self._read_normalizer_state("in", sd["input_normalizer"]),
                              "out": self._read_normalizer_state("out", sd["output_normalizer"])})
        self._ckpt_extra = {k: v for k, v in sd.items() if k == "optimizer"}

    def save(self, path):
        out = {"nn_state": {k: v.cpu() for k, v in self.state_dict().items()},
               "input_normalizer": self._normalizer_state("in"), "output_normalizer": self._normalizer_state("out")}
        out.update(self._ckpt_extra)
        torch.save(out, path)
