"""CPU oracle: an independent torch-CPU restatement of the CEE-US iCEM planning hot path.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this file; it is the *checker* for the CUDA path in
``cee-us_b200/`` and is never on the product path (the product raises if its CUDA library is missing).

Parity pin: every function below is checked against the committed fixtures in ``tests/golden/*.npz``
(``tests/test_oracle_golden.py``), which ``tests/golden/make_golden.py`` produced by running the reference's own,
unmodified code in the build container (``make_golden.py --check`` regenerates and compares them).  Upstream ships
no tests / golden vectors of its own.

Each function cites the reference ``file:line`` it restates (paths relative to ``/root/reference``).
All arithmetic is float32 torch on CPU, op order kept close to the reference so the timings taken from this
file by ``bench.py`` are representative of the reference's CPU torch path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np
import torch

# --------------------------------------------------------------------------------------------------
# Environment description (tensor-side constants of the two environments on the path)
# --------------------------------------------------------------------------------------------------


@dataclass
class EnvSpec:
    """Scalars the hot path reads from ``env`` (SURVEY.md 8b "Env hand-off").

    construction: mbrl/environments/fpp_construction_env.py:32-125
    playground:   mbrl/environments/playground_env_wgoals.py:24-72
    """

    kind: str  # "construction" | "playground"
    n_obj: int
    agent_dim: int
    dyn_dim: int
    stat_dim: int
    action_dim: int
    goal_dim: int
    action_low: np.ndarray
    action_high: np.ndarray
    case: str = "Singletower"
    sparse: bool = False
    shaped_reward: bool = True
    threshold: float = 0.02
    buffer_threshold: float = 0.025
    height_offset: float = 0.42  # Slide only (attribute of the simulator env, fpp_construction_env.py:101-105)

    @property
    def state_dim(self) -> int:  # observation without goal (obs_preproc)
        return self.agent_dim + self.n_obj * (self.dyn_dim + self.stat_dim)

    @property
    def obs_dim(self) -> int:
        return self.state_dim + self.goal_dim


def construction_spec(n_obj: int, case: str = "Singletower", sparse: bool = False, shaped_reward: bool = True) -> EnvSpec:
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
    return EnvSpec("construction", n_obj, 10, 12, 0, 4, 3 * n_obj + 3, -np.ones(4, np.float32), np.ones(4, np.float32),
                   case=case, sparse=sparse, shaped_reward=shaped_reward, threshold=thr, buffer_threshold=0.025)


def playground_spec(n_obj: int) -> EnvSpec:
    return EnvSpec("playground", n_obj, 4, 6, 3, 2, 2 * n_obj, -np.ones(2, np.float32), np.ones(2, np.float32),
                   case="playground", threshold=0.23, buffer_threshold=0.23)


# --------------------------------------------------------------------------------------------------
# Deterministic synthetic weights (shared by the golden generator, the tests and the bench)
# --------------------------------------------------------------------------------------------------


def _mlp_shapes(shapes, name, din, dout, hidden, num_layers, E, layer_norm=True):
    dims = [din] + [hidden] * num_layers
    for l in range(num_layers):
        shapes[f"{name}.layers.{l}.0.weight"] = (E, dims[l], dims[l + 1])
        shapes[f"{name}.layers.{l}.0.bias"] = (E, dims[l + 1])
        if layer_norm:
            shapes[f"{name}.layers.{l}.1.gamma"] = (E, dims[l + 1])
            shapes[f"{name}.layers.{l}.1.beta"] = (E, dims[l + 1])
    shapes[f"{name}.layers.{num_layers}.0.weight"] = (E, dims[-1], dout)
    shapes[f"{name}.layers.{num_layers}.0.bias"] = (E, dout)


def gnn_weight_shapes(spec: EnvSpec, hidden: int, num_layers: int, E: int, edge_global: bool = False) -> Dict[str, Tuple[int, ...]]:
    """state_dict layout of GraphNeuralNetworkEnsemble (gnn_modules.py:78-115, simple.py:36-57); edge_global ==
    ignore_global_v_node=False (:98-106)."""
    node_attr = spec.dyn_dim + spec.stat_dim
    shapes = {}
    _mlp_shapes(shapes, "edge_mlp", 2 * node_attr + spec.agent_dim + spec.action_dim, hidden, hidden, num_layers, E)
    _mlp_shapes(shapes, "node_mlp", node_attr + hidden + spec.action_dim + spec.agent_dim, spec.dyn_dim, hidden, num_layers, E)
    _mlp_shapes(shapes, "global_mlp", spec.agent_dim + hidden * (2 if edge_global else 1) + spec.action_dim, spec.agent_dim, hidden,
                num_layers, E)
    if edge_global:
        _mlp_shapes(shapes, "edge_mlp_global", node_attr + spec.agent_dim + spec.action_dim, hidden, hidden, num_layers, E)
    return shapes


def homog_weight_shapes(spec: EnvSpec, hidden: int, num_layers: int, E: int, emb: int) -> Dict[str, Tuple[int, ...]]:
    """state_dict layout of HomogeneousGraphNeuralNetworkEnsemble (gnn_modules.py:351-405)."""
    node_attr = spec.dyn_dim + spec.stat_dim
    shapes = {}
    _mlp_shapes(shapes, "embed_agent", spec.agent_dim, emb, hidden, 0, E)
    _mlp_shapes(shapes, "embed_obj", node_attr, emb, hidden, 0, E)
    _mlp_shapes(shapes, "output_head_agent", emb, spec.agent_dim, hidden, 0, E)
    _mlp_shapes(shapes, "output_head_objdyn", emb, spec.dyn_dim, hidden, 0, E)
    _mlp_shapes(shapes, "edge_mlp", 2 * emb + spec.action_dim, hidden, hidden, num_layers, E)
    _mlp_shapes(shapes, "node_mlp", emb + hidden + spec.action_dim, emb, hidden, num_layers, E)
    return shapes


def mlp_weight_shapes(spec: EnvSpec, hidden: int, num_layers: int, E: int) -> Dict[str, Tuple[int, ...]]:
    """state_dict layout of MLPParallelEnsemble for the dense model (mlp_ensemble.py:39-50, simple.py:36-57)."""
    dims = [spec.state_dim + spec.action_dim] + [hidden] * num_layers + [spec.state_dim]
    shapes = {}
    for l in range(num_layers + 1):
        shapes[f"layers.{l}.0.weight"] = (E, dims[l], dims[l + 1])
        shapes[f"layers.{l}.0.bias"] = (E, dims[l + 1])
    return shapes


def synth_weights(shapes: Dict[str, Tuple[int, ...]], seed: int, plain_init: bool = False) -> Dict[str, np.ndarray]:
    """Seeded weights.  Weight scale follows the reference initialiser N(0, (1/(2 sqrt(in)))^2)
    (torch_helpers.py:637-675; the "truncation" there is a no-op).  Unless ``plain_init``,
    biases / LN gamma / beta are randomised too so that parity tests exercise them (the reference's own
    init leaves them at 0 / 1 / 0)."""
    rs = np.random.RandomState(seed)
    out = {}
    for k in sorted(shapes):
        shp = shapes[k]
        if k.endswith("weight"):
            out[k] = (rs.standard_normal(shp) / (2.0 * math.sqrt(shp[1]))).astype(np.float32)
        elif k.endswith("bias"):
            out[k] = np.zeros(shp, np.float32) if plain_init else (0.1 * rs.standard_normal(shp)).astype(np.float32)
        elif k.endswith("gamma"):
            out[k] = np.ones(shp, np.float32) if plain_init else (1.0 + 0.1 * rs.standard_normal(shp)).astype(np.float32)
        elif k.endswith("beta"):
            out[k] = np.zeros(shp, np.float32) if plain_init else (0.1 * rs.standard_normal(shp)).astype(np.float32)
        else:
            raise KeyError(k)
    return out


def synth_normalizers(dims: Dict[str, int], seed: int, identity: bool) -> Dict[str, Tuple[np.ndarray, np.ndarray]]:
    """(mean, std) per normaliser group; identity == the state before the first train()."""
    rs = np.random.RandomState(seed + 7919)
    out = {}
    for k in sorted(dims):
        d = dims[k]
        if identity:
            out[k] = (np.zeros(d, np.float32), np.ones(d, np.float32))
        else:
            out[k] = ((0.1 * rs.standard_normal(d)).astype(np.float32), rs.uniform(0.5, 1.5, d).astype(np.float32))
    return out


def gnn_normalizer_dims(spec: EnvSpec) -> Dict[str, int]:
    return {"in_agent": spec.agent_dim, "in_dyn": spec.dyn_dim, "in_stat": spec.stat_dim, "in_action": spec.action_dim,
            "out_agent": spec.agent_dim, "out_dyn": spec.dyn_dim}


def mlp_normalizer_dims(spec: EnvSpec) -> Dict[str, int]:
    return {"in": spec.state_dim + spec.action_dim, "out": spec.state_dim}


def synth_start_obs(spec: EnvSpec, seed: int) -> Tuple[np.ndarray, np.ndarray]:
    """Synthetic start observation + goal (SURVEY.md 8d "Synthetic inputs")."""
    rs = np.random.RandomState(seed)
    N = spec.n_obj
    if spec.kind == "robodesk":
        obs = np.concatenate([0.3 * rs.standard_normal(spec.state_dim) + 0.5]).astype(np.float32)
        return obs, np.zeros(0, np.float32)
    if spec.kind == "construction":
        agent = np.concatenate([[1.34, 0.75, 0.53], rs.uniform(0, 0.05, 2), 0.01 * rs.standard_normal(5)])
        blocks = []
        for _ in range(N):
            pos = np.array([1.3, 0.75, 0.42]) + np.concatenate([rs.uniform(-0.15, 0.15, 2), [0.0]])
            blocks.append(np.concatenate([pos, 0.05 * rs.standard_normal(3), 0.01 * rs.standard_normal(6)]))
        goal = np.concatenate([np.array([1.3, 0.75, 0.42 + 0.05 * i]) for i in range(N)] + [np.zeros(3)])
        obs = np.concatenate([agent] + blocks + [goal])
    else:
        agent = np.concatenate([rs.uniform(-1.6, 1.6, 2), 0.1 * rs.standard_normal(2)])
        dyn = []
        for _ in range(N):
            r, th = rs.uniform(0.3, 1.4), rs.uniform(0, 2 * math.pi)
            dyn.append(np.concatenate([[r * math.cos(th), r * math.sin(th)], 0.05 * rs.standard_normal(4)]))
        stat = [rs.uniform(0, 1, 3) for _ in range(N)]
        goal = np.tile(rs.uniform(-1.7, 1.7, 2), N)
        obs = np.concatenate([agent] + dyn + stat + [goal])
    obs = obs.astype(np.float32)
    assert obs.shape[0] == spec.obs_dim
    return obs, obs[spec.state_dim:].copy()


# --------------------------------------------------------------------------------------------------
# Ensemble MLP (a9, a10)
# --------------------------------------------------------------------------------------------------

_ACTS: Dict[str, Optional[Callable]] = {
    "relu": torch.relu,
    "silu": torch.nn.functional.silu,
    "swish": lambda x: x * torch.sigmoid(x),
    "tanh": torch.tanh,
    "none": None,
}


def ensemble_layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """LayerNormEnsemble.forward, mbrl/models/utils.py:116-126 (biased variance, eps inside the sqrt)."""
    mean = x.mean(dim=-1, keepdim=True)
    var = torch.square(x - mean).mean(dim=-1, keepdim=True)
    y = (x - mean) / torch.sqrt(var + eps)
    return y * gamma.unsqueeze(1) + beta.unsqueeze(1)


def ensemble_mlp(x: torch.Tensor, w: Dict[str, torch.Tensor], prefix: str, num_layers: int, act: str,
                 layer_norm: bool, out_act: str = "none") -> torch.Tensor:
    """MLPParallelEnsemble.forward (simple.py:36-57, 84-86) with MLPParallelEnsembleLayer.forward
    (torch_parallel_ensembles/utils.py:26-32): y_e = x_e @ W_e + b_e for x [E, rows, in]."""
    f = _ACTS[act]
    for l in range(num_layers + 1):
        x = torch.bmm(x, w[f"{prefix}layers.{l}.0.weight"]) + w[f"{prefix}layers.{l}.0.bias"].unsqueeze(1)
        if l < num_layers:
            if layer_norm:
                x = ensemble_layernorm(x, w[f"{prefix}layers.{l}.1.gamma"], w[f"{prefix}layers.{l}.1.beta"])
            if f is not None:
                x = f(x)
    g = _ACTS[out_act]
    return g(x) if g is not None else x


# --------------------------------------------------------------------------------------------------
# Models
# --------------------------------------------------------------------------------------------------


def _t(a) -> torch.Tensor:
    return torch.as_tensor(np.asarray(a), dtype=torch.float32)


@dataclass
class GNNEnsembleOracle:
    """GNNForwardEnsembleModel inference half (gnn_ensemble.py:190-334, 876-983) around
    GraphNeuralNetworkEnsemble.forward (gnn_modules.py:173-278)."""

    spec: EnvSpec
    weights: Dict[str, np.ndarray]
    normalizers: Dict[str, Tuple[np.ndarray, np.ndarray]]
    hidden: int = 128
    num_layers: int = 2
    act: str = "relu"
    layer_norm: bool = True
    aggr: str = "mean"
    kind: str = "gnn"
    edge_global: bool = False  # ignore_agent_object: false (gnn_modules.py:98-106, 225-238)
    homogeneous: bool = False  # agent_as_global_node: false -> HomogeneousGraphNeuralNetworkEnsemble (:281-545)
    num_message_passing: int = 1
    _w: Dict[str, torch.Tensor] = field(default_factory=dict, repr=False)
    _edge_cache: Dict[int, tuple] = field(default_factory=dict, repr=False)

    def __post_init__(self):
        self._w = {k: _t(v) for k, v in self.weights.items()}
        self.E = next(iter(self._w.values())).shape[0]
        self._nrm = {k: (_t(m).view(1, -1), _t(s).view(1, -1)) for k, (m, s) in self.normalizers.items()}

    @property
    def ensemble_size(self) -> int:
        return self.E

    def _edges(self, B: int):
        # gnn_modules.py:148-170: nonzero() of (ones - eye) -> row-major (i asc, j asc, j != i); row = source i.
        if B not in self._edge_cache:
            N = self.spec.n_obj + (1 if self.homogeneous else 0)
            pairs = [(i, j) for i in range(N) for j in range(N) if i != j]
            base = torch.tensor(pairs, dtype=torch.int64)
            off = (torch.arange(B, dtype=torch.int64) * N).repeat_interleave(len(pairs)).unsqueeze(-1)
            el = base.repeat(B, 1) + off
            row, col = el[:, 0].contiguous(), el[:, 1].contiguous()
            edge_batch = torch.div(row, N, rounding_mode="trunc")
            node_batch = torch.arange(B, dtype=torch.int64).repeat_interleave(N)
            self._edge_cache = {B: (row, col, edge_batch, node_batch)}
        return self._edge_cache[B]

    def _segment(self, vals: torch.Tensor, ids: torch.Tensor, num: int) -> torch.Tensor:
        # models/utils.py:30-51
        E, _, Hd = vals.shape
        idx = ids.view(1, -1, 1).expand(E, -1, Hd)
        out = vals.new_zeros((E, num, Hd)).scatter_add_(1, idx, vals)
        if self.aggr == "mean":
            cnt = vals.new_zeros((E, num, Hd)).scatter_add_(1, idx, torch.ones_like(vals))
            out = out / cnt.clamp(1.0)
        return out

    def _homog_forward(self, agent, dyn, stat, action):
        """HomogeneousGraphNeuralNetworkEnsemble.forward (gnn_modules.py:499-545)."""
        E, B, N, _ = dyn.shape
        M = N + 1
        w = self._w
        lin = lambda x, p: torch.bmm(x, w[f"{p}.layers.0.0.weight"]) + w[f"{p}.layers.0.0.bias"].unsqueeze(1)
        obj = torch.cat([dyn, stat], dim=-1) if self.spec.stat_dim > 0 else dyn
        f_agent = lin(agent, "embed_agent")                                   # [E,B,F]
        f_obj = lin(obj.reshape(E, B * N, -1), "embed_obj").view(E, B, N, -1)
        feat = torch.cat([f_agent.unsqueeze(2), f_obj], dim=2)                # agent = node 0
        row, col, edge_batch, node_batch = self._edges(B)
        for _ in range(self.num_message_passing):                             # :449-497
            node = feat.reshape(E, B * M, -1)
            edge_in = torch.cat([node[:, row], node[:, col], action[:, edge_batch]], dim=-1)
            msg = ensemble_mlp(edge_in, w, "edge_mlp.", self.num_layers, self.act, self.layer_norm)
            agg = self._segment(msg, row, B * M)
            node_in = torch.cat([node, action[:, node_batch], agg], dim=-1)
            feat = ensemble_mlp(node_in, w, "node_mlp.", self.num_layers, self.act, self.layer_norm).view(E, B, M, -1)
        d_agent = lin(feat[:, :, 0].contiguous(), "output_head_agent")
        d_dyn = lin(feat[:, :, 1:].reshape(E, B * N, -1), "output_head_objdyn").view(E, B, N, -1)
        return d_agent, d_dyn

    def gnn_forward(self, agent, dyn, stat, action):
        """[E,B,A], [E,B,N,D], [E,B,N,S] (S may be 0), [E,B,U] -> (d_agent [E,B,A], d_dyn [E,B,N,D]); normalised."""
        if self.homogeneous:
            return self._homog_forward(agent, dyn, stat, action)
        E, B, N, _ = dyn.shape
        node = torch.cat([dyn, stat], dim=-1) if self.spec.stat_dim > 0 else dyn
        node = node.reshape(E, B * N, -1)
        row, col, edge_batch, node_batch = self._edges(B)
        ctx = torch.cat([agent, action], dim=-1)  # [E,B,A+U]
        edge_in = torch.cat([node[:, row], node[:, col], ctx[:, edge_batch]], dim=-1)
        msg = ensemble_mlp(edge_in, self._w, "edge_mlp.", self.num_layers, self.act, self.layer_norm)
        agg_node = self._segment(msg, row, B * N)
        node_in = torch.cat([node, ctx[:, node_batch], agg_node], dim=-1)
        d_dyn = ensemble_mlp(node_in, self._w, "node_mlp.", self.num_layers, self.act, self.layer_norm)
        agg_glob = self._segment(msg, edge_batch, B)
        glob_in = ctx
        if self.edge_global:  # :225-238, 244-246
            ng = ensemble_mlp(torch.cat([node, ctx[:, node_batch]], dim=-1), self._w, "edge_mlp_global.", self.num_layers,
                              self.act, self.layer_norm)
            glob_in = torch.cat([ctx, self._segment(ng, node_batch, B)], dim=-1)
        glob_in = torch.cat([glob_in, agg_glob], dim=-1)
        d_agent = ensemble_mlp(glob_in, self._w, "global_mlp.", self.num_layers, self.act, self.layer_norm)
        return d_agent, d_dyn.view(E, B, N, -1)

    def predict(self, state: torch.Tensor, action: torch.Tensor) -> torch.Tensor:
        """One model step on goal-free states: [E,B,state_dim], [E,B,U] -> next state [E,B,state_dim].
        gnn_ensemble.py:903-944 without obs_preproc/obs_postproc (the goal is handled by the caller)."""
        s = self.spec
        E, B, _ = state.shape
        A, D, S, N = s.agent_dim, s.dyn_dim, s.stat_dim, s.n_obj
        agent_raw = state[..., :A]
        dyn_raw = state[..., A:A + N * D].reshape(E, B, N, D)
        stat_raw = state[..., A + N * D:].reshape(E, B, N, S)
        nm = self._nrm
        agent = (agent_raw - nm["in_agent"][0]) / nm["in_agent"][1]
        dyn = (dyn_raw - nm["in_dyn"][0]) / nm["in_dyn"][1]
        stat = (stat_raw - nm["in_stat"][0]) / nm["in_stat"][1] if S > 0 else stat_raw
        act = (action - nm["in_action"][0]) / nm["in_action"][1]
        d_agent, d_dyn = self.gnn_forward(agent, dyn, stat, act)
        # _obs_invert_preprocessing (gnn_ensemble.py:302-334): output normalisers for the deltas, the *input*
        # static normaliser for the normalise->denormalise round trip of the static features (:320-324)
        agent_out = nm["out_agent"][1] * d_agent + nm["out_agent"][0]
        dyn_out = (nm["out_dyn"][1] * d_dyn + nm["out_dyn"][0]).reshape(E, B, N * D)
        parts = [agent_out + agent_raw, dyn_out + dyn_raw.reshape(E, B, N * D)]
        if S > 0:
            parts.append((nm["in_stat"][1] * stat + nm["in_stat"][0]).reshape(E, B, N * S))
        return torch.cat(parts, dim=-1)


@dataclass
class MLPEnsembleOracle:
    """ParallelNNDeterministicEnsemble.predict, mlp_ensemble.py:530-573."""

    spec: EnvSpec
    weights: Dict[str, np.ndarray]
    normalizers: Dict[str, Tuple[np.ndarray, np.ndarray]]
    hidden: int = 256
    num_layers: int = 3
    act: str = "silu"
    layer_norm: bool = False
    kind: str = "mlp"
    _w: Dict[str, torch.Tensor] = field(default_factory=dict, repr=False)

    def __post_init__(self):
        self._w = {k: _t(v) for k, v in self.weights.items()}
        self.E = next(iter(self._w.values())).shape[0]
        self._nrm = {k: (_t(m).view(1, -1), _t(s).view(1, -1)) for k, (m, s) in self.normalizers.items()}

    @property
    def ensemble_size(self) -> int:
        return self.E

    def predict(self, state: torch.Tensor, action: torch.Tensor) -> torch.Tensor:
        x = (torch.cat([state, action], dim=-1) - self._nrm["in"][0]) / self._nrm["in"][1]
        y = ensemble_mlp(x, self._w, "", self.num_layers, self.act, self.layer_norm)
        return state + (y * self._nrm["out"][1] + self._nrm["out"][0])


def rollout(model, start_obs: torch.Tensor, actions: torch.Tensor, goal: torch.Tensor) -> torch.Tensor:
    """predict_n_steps / rollout_generator (gnn_ensemble.py:876-901,946-983; mlp_ensemble.py:502-528,575-611).

    start_obs [P,O] (with its own goal dims), actions [P,H,U], goal [G] (= env.goal, appended to every
    *predicted* observation by env.obs_postproc, fpp_construction_env.py:133-138).
    Returns traj [H+1,E,P,O]: traj[h] = observations at step h, traj[h+1] = next_observations at step h
    (the reference's all_observations_ / all_next_observations_ buffers are traj[:-1] / traj[1:])."""
    P, H, U = actions.shape
    E, sd = model.ensemble_size, model.spec.state_dim
    O = model.spec.obs_dim
    traj = torch.empty((H + 1, E, P, O), dtype=torch.float32)
    traj[0] = start_obs.unsqueeze(0).expand(E, -1, -1)
    state = traj[0, :, :, :sd]
    g = goal.view(1, 1, -1).expand(E, P, -1)
    for h in range(H):
        a = actions[:, h].unsqueeze(0).expand(E, -1, -1)
        state = model.predict(state, a)
        traj[h + 1] = torch.cat([state, g], dim=-1)
    return traj


# --------------------------------------------------------------------------------------------------
# Costs (a14-a17)
# --------------------------------------------------------------------------------------------------


def disagreement_bonus(next_obs_pehd: torch.Tensor) -> torch.Tensor:
    """TorchCuriosityMpcICem._model_epistemic_costs (mpc_torch_curiosity.py:45-51): unbiased std over the
    ensemble of every observation dim, summed over dims.  [P,E,H,O] -> [P,H]."""
    return torch.std(next_obs_pehd, dim=1).sum(dim=-1)


def construction_cost(spec: EnvSpec, obs: torch.Tensor, next_obs: torch.Tensor) -> torch.Tensor:
    """FetchPickAndPlaceConstruction.cost_fn (fpp_construction_env.py:404-511).
    obs, next_obs: [P,E,H,O] -> [P,E,H]."""
    N, A, D = spec.n_obj, spec.agent_dim, spec.dyn_dim
    sd = spec.state_dim
    if spec.case in ("Flip", "Slide"):
        # per-coordinate distances (:297-355, :422-439): the achieved goal is the euler-angle triple of every block for Flip
        # (:42-44), its position for Slide; parameters :93-105
        flip = spec.case == "Flip"
        ao = 3 if flip else 0
        dev = next_obs.device
        w = torch.tensor([1.0, 0.0, 0.0] if flip else [1.0, 1.0, 1.0], device=dev)
        buf = torch.tensor(1e-4, device=dev) if flip else torch.tensor([0.1, 0.1, spec.height_offset], device=dev)
        thres = w * 0.087 if flip else torch.tensor([0.1, 0.1, spec.height_offset], device=dev)
        goal = next_obs[..., sd:sd + 3 * N + 3]
        per = [torch.maximum(torch.abs(next_obs[..., A + D * i + ao:A + D * i + ao + 3] - goal[..., 3 * i:3 * i + 3]), buf) * w
               for i in range(N)]
        sparse_obj = torch.stack([torch.any(p > thres, dim=-1).float() for p in per])          # :429-432
        exp_dense = torch.stack([torch.mean(1.0 - torch.exp(-0.5 * p), dim=-1) for p in per])   # :323-338, :434-439
        g = torch.zeros(next_obs.shape[:-1], dtype=torch.float32, device=dev)
        if spec.shaped_reward:
            assert not spec.case == "Slide", "Shaped reward cannot be used with Slide!"          # :455
            g = torch.linalg.norm(next_obs[..., 0:3] - torch.tensor([1.344, 0.749, 0.5319], device=dev), dim=-1)  # :457-458, :119
        if spec.sparse:
            return sparse_obj.sum(dim=0) + g * 0.001                                              # :476-480
        return exp_dense.sum(dim=0) * 1e-3 + sparse_obj.sum(dim=0) + g * 0.01                     # :498-503

    def goal_of(o):
        return o[..., sd:sd + 3 * N + 3]

    def achieved_of(o):  # achieved_goal_idx: block positions then gripper position (:41-53)
        return torch.cat([o[..., A + D * i:A + D * i + 3] for i in range(N)] + [o[..., 0:3]], dim=-1)

    def dists(ach, goal):  # subgoal_distances :282-296
        return torch.stack([torch.linalg.norm(ach[..., 3 * i:3 * i + 3] - goal[..., 3 * i:3 * i + 3], dim=-1)
                            for i in range(N)])

    goal, ach = goal_of(next_obs), achieved_of(next_obs)
    d = dists(ach, goal)  # [N,P,E,H]
    grip = ach[..., -3:]
    block_pos = torch.cat([ach[..., :-3], goal[..., -3:]], dim=-1)
    g = torch.zeros(next_obs.shape[:-1], dtype=torch.float32, device=next_obs.device)
    if spec.shaped_reward:
        unsolved = (d > spec.threshold).float()
        start = obs[..., 0:1, :]
        d0 = dists(achieved_of(start), goal_of(start))
        k = int(N - (d0 > spec.threshold).float()[..., 0].sum(dim=0).view(-1)[0].item())  # get_next_block_id :365-383
        g = torch.linalg.norm(grip - block_pos[..., 3 * k:3 * k + 3], dim=-1)
        nxt_all = N - unsolved.sum(dim=0)
        g = g * (nxt_all <= k).float()
    if spec.sparse:
        return (d > spec.threshold).float().sum(dim=0) + (g > spec.threshold).float() * 0.5
    if "tower" in spec.case or spec.case == "Pyramid":
        return (d > spec.threshold).float().sum(dim=0) + g * 0.01
    return torch.maximum(d, torch.tensor(spec.buffer_threshold, device=next_obs.device)).sum(dim=0) + g * 0.01


def playground_cost(spec: EnvSpec, next_obs: torch.Tensor) -> torch.Tensor:
    """PlaygroundwGoals.cost_fn/_compute_reward (playground_env_wgoals.py:108-157). [P,E,H,O] -> [P,E,H]."""
    N, A, D = spec.n_obj, spec.agent_dim, spec.dyn_dim
    sd = spec.state_dim
    goal = next_obs[..., sd:sd + 2 * N]
    ach = torch.stack([next_obs[..., A + D * i + c] for i in range(N) for c in (0, 1)], dim=-1)
    diff = ach - goal
    return torch.norm(torch.maximum(torch.abs(diff), torch.tensor(0.23)), dim=-1)


def robodesk_spec(task: str = "open_slide", reward: str = "dense") -> EnvSpec:
    """Robodesk (robodesk_env.py:14-31): agent 24, 8 objects x (13 dynamic + 6 static), 8 action dims, no goal."""
    return EnvSpec("robodesk", 8, 24, 13, 6, 8, 0, -np.ones(8, np.float32), np.ones(8, np.float32), case=task,
                   sparse=reward == "sparse", shaped_reward=False, threshold=0.0, buffer_threshold=0.0)


def robodesk_cost(spec: EnvSpec, next_obs: torch.Tensor) -> torch.Tensor:
    """Robodesk.cost_fn = -reward (robodesk_env.py:53-135).  [...,O] -> [...]."""
    o, task, sparse = next_obs, spec.case, spec.sparse
    dist = lambda t: torch.linalg.norm(o[..., :3] - t, dim=-1)
    if task == "open_slide":
        r = (o[..., 37] + 0.3) - 0.55
        r = (r >= 0.0).float() if sparse else r - 0.05 * dist(o[..., 37:40])
    elif task == "open_drawer":
        r = -0.2 - (o[..., 25] - 0.85)
        r = (r >= 0.0).float() if sparse else r - 0.05 * dist(o[..., 24:27])
    elif "push" in task:
        zi = {"push_red": 52, "push_green": 65, "push_blue": 78}[task]
        r = -0.0005238 - (o[..., zi] - 0.76)
        if sparse:
            r = (r >= 0.001).float()
        else:
            tgt = o[..., zi - 2:zi + 1].clone()
            tgt[..., -1] = 0.76
            r = r - 0.05 * dist(tgt)
    else:
        name = task[:-len("_off_table")]
        bx = {"ball": 89, "upright_block": 102, "flat_block": 115}[name]
        z0 = {"ball": 0.79963282, "upright_block": 0.84978449, "flat_block": 0.77478449}[name]
        r = (o[..., bx + 2] < 0.6).float() if sparse else 1 - (o[..., bx + 2] / z0) - 0.001 * dist(o[..., bx:bx + 3])
    return -r


def env_cost(spec: EnvSpec, obs: torch.Tensor, next_obs: torch.Tensor) -> torch.Tensor:
    if spec.kind == "robodesk":
        return robodesk_cost(spec, next_obs)
    return construction_cost(spec, obs, next_obs) if spec.kind == "construction" else playground_cost(spec, next_obs)


# --------------------------------------------------------------------------------------------------
# Colored noise (a2)
# --------------------------------------------------------------------------------------------------


def powerlaw_scales(beta: float, H: int) -> Tuple[np.ndarray, float]:
    """s_k and sigma of torch_powerlaw_psd_gaussian (colored_noise.py:174-187)."""
    f = np.fft.rfftfreq(H).astype(np.float32)
    fmin = max(0.0, 1.0 / H)
    ix = int(np.sum(f < fmin))
    if ix and ix < len(f):
        f[:ix] = f[ix]
    s = f ** np.float32(-beta / 2.0)
    w = s[1:].copy()
    w[-1] *= (1 + (H % 2)) / 2.0
    sigma = 2.0 * math.sqrt(float(np.sum(w.astype(np.float64) ** 2))) / H
    return s.astype(np.float32), float(sigma)


def colored_noise_from_gaussians(zr: torch.Tensor, zi: torch.Tensor, beta: float, H: int) -> torch.Tensor:
    """y = irfft((sr + i si)) / sigma with sr = s_k zr, si = s_k zi and the DC / Nyquist fixes
    (colored_noise.py:197-216).  zr, zi: standard normal [..., H//2+1] -> [..., H]."""
    s, sigma = powerlaw_scales(beta, H)
    st = torch.from_numpy(s)
    sr, si = zr * st, zi * st
    sr, si = sr.clone(), si.clone()
    if H % 2 == 0:
        si[..., -1] = 0
        sr[..., -1] *= math.sqrt(2)
    si[..., 0] = 0
    sr[..., 0] *= math.sqrt(2)
    return torch.fft.irfft(torch.complex(sr, si), n=H, dim=-1) / sigma


def irdft_matrix(beta: float, H: int) -> np.ndarray:
    """The same transform as a fixed real matrix M [2F, H]: y = [zr | zi] @ M.
    This is what the device sampler multiplies by; built in float64, returned as float32."""
    F = H // 2 + 1
    s, sigma = powerlaw_scales(beta, H)
    s = s.astype(np.float64)
    M = np.zeros((2 * F, H), np.float64)
    t = np.arange(H)
    for k in range(F):
        nyq = (H % 2 == 0) and (k == F - 1)
        if k == 0:
            M[k] = s[0] * math.sqrt(2)
        elif nyq:
            M[k] = s[k] * math.sqrt(2) * np.cos(math.pi * t)
        else:
            M[k] = 2 * s[k] * np.cos(2 * math.pi * k * t / H)
            M[F + k] = -2 * s[k] * np.sin(2 * math.pi * k * t / H)
    return (M / (H * sigma)).astype(np.float32)


# --------------------------------------------------------------------------------------------------
# Philox4x32-10 + Box-Muller: bit-level restatement of the device sampler's RNG (cee-us_b200/csrc)
# --------------------------------------------------------------------------------------------------

_PHILOX_M0, _PHILOX_M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_PHILOX_W0, _PHILOX_W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox4x32(counter: np.ndarray, key: np.ndarray) -> np.ndarray:
    """Philox4x32-10 (Salmon et al. 2011, Random123).  counter [n,4] uint32, key [n,2] uint32 -> [n,4] uint32."""
    c = counter.astype(np.uint32).copy()
    k0, k1 = key[:, 0].astype(np.uint32).copy(), key[:, 1].astype(np.uint32).copy()
    mask = np.uint64(0xFFFFFFFF)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _PHILOX_M0 * c[:, 0].astype(np.uint64)
            p1 = _PHILOX_M1 * c[:, 2].astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & mask).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & mask).astype(np.uint32)
            c = np.stack([hi1 ^ c[:, 1] ^ k0, lo1, hi0 ^ c[:, 3] ^ k1, lo0], axis=1)
            k0 = k0 + _PHILOX_W0
            k1 = k1 + _PHILOX_W1
    return c


def philox_normals(seed: int, stream: int, n_rows: int, n_per_row: int) -> np.ndarray:
    """Standard normals exactly as the device sampler draws them: for row r (a (trajectory, action-dim)
    pair, *global* index so results do not depend on the GPU count) and block b of 4 values,
    counter = (b, r_lo, r_hi, stream), key = (seed_lo, seed_hi); Box-Muller on (u0,u1) and (u2,u3)
    with u = (x + 0.5) * 2^-32."""
    nb = (n_per_row + 3) // 4
    r = np.repeat(np.arange(n_rows, dtype=np.uint64), nb)
    b = np.tile(np.arange(nb, dtype=np.uint64), n_rows)
    ctr = np.stack([b & np.uint64(0xFFFFFFFF), r & np.uint64(0xFFFFFFFF), r >> np.uint64(32),
                    np.full_like(r, np.uint64(stream & 0xFFFFFFFF))], axis=1).astype(np.uint32)
    key = np.tile(np.array([[seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF]], dtype=np.uint32), (len(r), 1))
    x = philox4x32(ctr, key)
    u = (x.astype(np.float64) + 0.5) * (2.0 ** -32)
    u = u.astype(np.float32)
    rad0 = np.sqrt(np.float32(-2.0) * np.log(u[:, 0]))
    rad1 = np.sqrt(np.float32(-2.0) * np.log(u[:, 2]))
    two_pi = np.float32(2.0 * math.pi)
    z = np.stack([rad0 * np.cos(two_pi * u[:, 1]), rad0 * np.sin(two_pi * u[:, 1]),
                  rad1 * np.cos(two_pi * u[:, 3]), rad1 * np.sin(two_pi * u[:, 3])], axis=1).astype(np.float32)
    return z.reshape(n_rows, nb * 4)[:, :n_per_row]


# --------------------------------------------------------------------------------------------------
# iCEM controller (a1-a5, a14, a15, a18-a20)
# --------------------------------------------------------------------------------------------------


@dataclass
class ICEMParams:
    """controller_params / action_sampler_params of the settings yamls."""

    horizon: int = 30
    num_traj: int = 128
    opt_iterations: int = 3
    elites_size: int = 10
    alpha: float = 0.1
    init_std: float = 0.8
    relative_init: bool = True
    execute_best_elite: bool = True
    keep_previous_elites: bool = True
    shift_elites_over_time: bool = True
    fraction_elites_reused: float = 0.3
    use_mean_actions: bool = True
    colored_noise: bool = True
    noise_beta: float = 3.5
    use_ensemble_cost_std: bool = False
    cost_along_trajectory: str = "sum"  # "sum" | "best"
    curiosity: bool = True  # TorchCuriosityMpcICem vs TorchMpcICem
    extrinsic_reward: bool = False
    extrinsic_reward_scale: float = 1.0
    stable_sort: bool = True  # fully_deterministic (mpc_torch.py:451-456)

    @property
    def num_elites(self) -> int:  # _check_validity_parameters, mpc_torch.py:514-519
        return max(2, min(self.elites_size, self.num_traj // 2))

    @property
    def num_kept(self) -> int:  # int(len(elites) * fraction), mpc_torch.py:262,321
        return int(self.num_elites * self.fraction_elites_reused)


class ICEMOracle:
    """TorchMpcICem / TorchCuriosityMpcICem.get_action (mpc_torch.py:282-439, mpc_torch_curiosity.py:53-80).

    ``noise_fn(size)`` stands for colored_noise.torch_powerlaw_psd_gaussian(beta, size) -- the RNG boundary
    (SURVEY.md 8c); it is called with sizes (P,U,H) and (n_kept,U,H) in the reference's order."""

    def __init__(self, model, params: ICEMParams, noise_fn: Callable[[Tuple[int, int, int]], torch.Tensor]):
        self.model, self.p, self.noise_fn = model, params, noise_fn
        self.spec: EnvSpec = model.spec
        self.low, self.high = _t(self.spec.action_low), _t(self.spec.action_high)
        H, U = params.horizon, self.spec.action_dim
        self.mean_ = torch.zeros(H, U)
        self.std_ = torch.ones(H, U)
        self.elite_actions: Optional[torch.Tensor] = None  # [k,H,U]
        self.elite_traj: Optional[torch.Tensor] = None  # [H+1,E,k,O]
        self.goal = torch.zeros(self.spec.goal_dim)
        self.trace: List[dict] = []

    # mpc_torch.py:159-202
    def beginning_of_rollout(self):
        if self.p.relative_init:
            self.mean_ = ((self.high + self.low) * 2.0).expand_as(self.mean_).clone()
        else:
            self.mean_.zero_()
        self._reset_std()
        self.elite_actions = self.elite_traj = None

    def _reset_std(self):
        H = self.p.horizon
        if self.p.relative_init:
            self.std_ = ((self.high - self.low) * (self.p.init_std / 2.0)).expand(H, -1).clone()
        else:
            self.std_ = torch.full_like(self.std_, self.p.init_std)

    def set_goal(self, goal):
        self.goal = _t(goal)

    # mpc_torch.py:204-236
    def _sample(self, num: int) -> torch.Tensor:
        H, U = self.p.horizon, self.spec.action_dim
        y = self.noise_fn((num, U, H)).transpose(2, 1)  # [num,H,U]
        s = y * self.std_ + self.mean_
        s = torch.max(torch.min(s, self.high), self.low)
        return s.clone()

    def _costs_per_model(self, traj: torch.Tensor) -> torch.Tensor:
        """traj [H+1,E,P',O] -> costs_per_model [P',E]."""
        p = self.p
        obs = traj[:-1].permute(2, 1, 0, 3)
        nxt = traj[1:].permute(2, 1, 0, 3)
        E = traj.shape[1]
        if p.curiosity:
            path = -disagreement_bonus(nxt).unsqueeze(1).expand(-1, E, -1)
            if p.extrinsic_reward:
                path = path + p.extrinsic_reward_scale * env_cost(self.spec, obs, nxt)
        else:
            path = env_cost(self.spec, obs, nxt)
        if p.cost_along_trajectory == "sum":
            return path.sum(dim=-1)
        if p.cost_along_trajectory == "best":
            return torch.amin(path[..., 1:], dim=-1)
        raise NotImplementedError(p.cost_along_trajectory)

    def get_action(self, obs: np.ndarray) -> np.ndarray:
        p = self.p
        P, H = p.num_traj, p.horizon
        start = _t(obs).view(1, -1).expand(P, -1)
        step_trace = {"iters": []}
        for i in range(p.opt_iterations):
            acts = self._sample(P)
            if p.use_mean_actions and i == p.opt_iterations - 1:
                acts[0] = self.mean_
            if i == 0 and p.shift_elites_over_time and self.elite_actions is not None:  # :249-267, :304-310
                n = p.num_kept
                reused = self.elite_actions[:n, 1:]
                last = self._sample(n)[:, -1:]
                acts[:n] = torch.cat([reused, last], dim=1)
            traj = rollout(self.model, start, acts, self.goal)
            if i > 0 and p.keep_previous_elites:  # :319-321
                n = p.num_kept
                traj = torch.cat([traj, self.elite_traj[:, :, :n]], dim=2)
                acts = torch.cat([acts, self.elite_actions[:n]], dim=0)
            cpm = self._costs_per_model(traj)
            costs = cpm.mean(dim=-1)
            costs_std = cpm.std(dim=-1)
            if p.use_ensemble_cost_std:
                costs = costs + costs_std
            best = int(torch.argmin(costs))
            # update_distributions :445-478
            order = torch.sort(costs, stable=True)[1] if p.stable_sort else torch.argsort(costs)
            eidx = order[: p.num_elites]
            self.elite_actions = acts[eidx].clone()
            self.elite_traj = traj[:, :, eidx].clone()
            E = traj.shape[1]
            es = self.elite_actions.unsqueeze(1).expand(-1, E, -1, -1)
            new_mean = es.mean(dim=(0, 1))
            new_std = es.std(dim=(0, 1))
            self.mean_ = p.alpha * self.mean_ + (1 - p.alpha) * new_mean
            self.std_ = p.alpha * self.std_ + (1 - p.alpha) * new_std
            step_trace["iters"].append(dict(costs=costs.clone(), costs_per_model=cpm.clone(), elite_idx=eidx.clone(),
                                            best=best, mean=self.mean_.clone(), std=self.std_.clone(),
                                            actions=acts.clone()))
        if p.execute_best_elite:
            action = acts[best][0].clone()
        else:
            action = self.mean_[0].clone()
        self.mean_[:-1] = self.mean_[1:].clone()  # :413-417 (last entry kept)
        self._reset_std()
        step_trace["action"] = action.numpy().copy()
        step_trace["best_cost"] = float(costs.min())
        self.trace.append(step_trace)
        return action.numpy()


def flops_per_model_step(model) -> int:
    """Algorithmic GEMM FLOPs (MACs x 2) per (trajectory, ensemble member, horizon step), SURVEY.md 8d."""
    s = model.spec
    Hd, L = model.hidden, model.num_layers
    if model.kind == "mlp":
        dims = [s.state_dim + s.action_dim] + [Hd] * L + [s.state_dim]
        return 2 * sum(a * b for a, b in zip(dims[:-1], dims[1:]))
    na = s.dyn_dim + s.stat_dim
    N = s.n_obj

    def mlp(din, dout):
        dims = [din] + [Hd] * L + [dout]
        return sum(a * b for a, b in zip(dims[:-1], dims[1:]))

    e = mlp(2 * na + s.agent_dim + s.action_dim, Hd)
    n = mlp(na + s.agent_dim + s.action_dim + Hd, s.dyn_dim)
    g = mlp(s.agent_dim + s.action_dim + Hd, s.agent_dim)
    return 2 * (N * (N - 1) * e + N * n + g)
