"""Batched training of the dynamics ensembles on the GPU: the ``train()`` half of

  * ``GNNForwardEnsembleModel``          mbrl/models/gnn_ensemble.py:336-720
  * ``ParallelNNDeterministicEnsemble``  mbrl/models/mlp_ensemble.py:172-429

with the reference's semantics: normaliser update from the rollout buffer (``Normalizer_v2`` or the running ``Normalizer``,
mbrl/torch_helpers.py:767-883), per-member batches drawn by ``TorchTrainingIterator`` (bootstrapped / non-bootstrapped epochs or a
fixed number of iterations, torch_helpers.py:344-445 -- the same ``np.random`` calls in the same order, so a seeded run sees the same
batches as the reference), loss = MSE averaged over batch and output dims and SUMMED over the members, Adam with ``eps=1e-4`` and
weight decay, optional keeping of the best validation weights.

Training is not the planning hot path: the forward / backward passes here are torch autograd over cuBLAS batched GEMMs on the
CUDA device (library code), all E members in one ``bmm``.  After ``train()`` the planner's engines re-upload the weights on their
next call (``weights_version``).
"""
import numpy as np
import torch
import torch.nn.functional as F

_ACT = {"relu": torch.relu, "silu": F.silu, "swish": F.silu, "tanh": torch.tanh, "none": None, None: None}


def ensemble_mlp(x, w, prefix, num_layers, act, layer_norm):
    """MLPParallelEnsemble.forward (torch_parallel_ensembles/simple.py:36-57): x [E,rows,in] -> [E,rows,out]."""
    f = _ACT[act]
    for l in range(num_layers + 1):
        x = torch.baddbmm(w[f"{prefix}layers.{l}.0.bias"].unsqueeze(1), x, w[f"{prefix}layers.{l}.0.weight"])
        if l < num_layers:
            if layer_norm:  # LayerNormEnsemble, models/utils.py:116-126 (biased variance, eps inside the sqrt)
                mean = x.mean(dim=-1, keepdim=True)
                var = torch.square(x - mean).mean(dim=-1, keepdim=True)
                x = (x - mean) / torch.sqrt(var + 1e-5)
                x = x * w[f"{prefix}layers.{l}.1.gamma"].unsqueeze(1) + w[f"{prefix}layers.{l}.1.beta"].unsqueeze(1)
            if f is not None:
                x = f(x)
    return x


def _segment(vals, ids, num, mean):
    """unsorted_segment_{mean,sum}_ensemble (models/utils.py:30-51)."""
    E, _, Hd = vals.shape
    idx = ids.view(1, -1, 1).expand(E, -1, Hd)
    out = vals.new_zeros((E, num, Hd)).scatter_add(1, idx, vals)
    if mean:
        cnt = vals.new_zeros((E, num, Hd)).scatter_add(1, idx, torch.ones_like(vals))
        out = out / cnt.clamp(1.0)
    return out


def _edges(B, M, device):
    pairs = torch.tensor([(i, j) for i in range(M) for j in range(M) if i != j], dtype=torch.int64, device=device)
    off = (torch.arange(B, dtype=torch.int64, device=device) * M).repeat_interleave(pairs.shape[0]).unsqueeze(-1)
    el = pairs.repeat(B, 1) + off
    row, col = el[:, 0].contiguous(), el[:, 1].contiguous()
    return row, col, torch.div(row, M, rounding_mode="trunc"), torch.arange(B, dtype=torch.int64, device=device).repeat_interleave(M)


def gnn_forward(w, m, agent, dyn, stat, action):
    """GraphNeuralNetworkEnsemble / HomogeneousGraphNeuralNetworkEnsemble .forward (gnn_modules.py:173-278, 449-545) on normalised
    inputs [E,B,A], [E,B,N,D], [E,B,N,S], [E,B,U] -> (d_agent [E,B,A], d_dyn [E,B,N,D])."""
    E, B, N, _ = dyn.shape
    L, act, ln, mean = m.num_layers, m.act_fn, m.layer_norm, m.aggr_fn == "mean"
    obj = torch.cat([dyn, stat], dim=-1) if stat.shape[-1] > 0 else dyn
    if not m.agent_as_global_node:
        M = N + 1
        lin = lambda x, p: torch.baddbmm(w[f"{p}.layers.0.0.bias"].unsqueeze(1), x, w[f"{p}.layers.0.0.weight"])
        feat = torch.cat([lin(agent, "embed_agent").unsqueeze(2), lin(obj.reshape(E, B * N, -1), "embed_obj").view(E, B, N, -1)], dim=2)
        row, col, edge_batch, node_batch = _edges(B, M, dyn.device)
        for _ in range(m.num_message_passing):
            node = feat.reshape(E, B * M, -1)
            msg = ensemble_mlp(torch.cat([node[:, row], node[:, col], action[:, edge_batch]], dim=-1), w, "edge_mlp.", L, act, ln)
            node_in = torch.cat([node, action[:, node_batch], _segment(msg, row, B * M, mean)], dim=-1)
            feat = ensemble_mlp(node_in, w, "node_mlp.", L, act, ln).view(E, B, M, -1)
        return (lin(feat[:, :, 0].contiguous(), "output_head_agent"),
                lin(feat[:, :, 1:].reshape(E, B * N, -1), "output_head_objdyn").view(E, B, N, -1))
    node = obj.reshape(E, B * N, -1)
    row, col, edge_batch, node_batch = _edges(B, N, dyn.device)
    ctx = torch.cat([agent, action], dim=-1)
    msg = ensemble_mlp(torch.cat([node[:, row], node[:, col], ctx[:, edge_batch]], dim=-1), w, "edge_mlp.", L, act, ln)
    node_in = torch.cat([node, ctx[:, node_batch], _segment(msg, row, B * N, mean)], dim=-1)
    d_dyn = ensemble_mlp(node_in, w, "node_mlp.", L, act, ln)
    glob_in = ctx
    if not m.ignore_agent_object:
        ng = ensemble_mlp(torch.cat([node, ctx[:, node_batch]], dim=-1), w, "edge_mlp_global.", L, act, ln)
        glob_in = torch.cat([ctx, _segment(ng, node_batch, B, mean)], dim=-1)
    glob_in = torch.cat([glob_in, _segment(msg, edge_batch, B, mean)], dim=-1)
    return ensemble_mlp(glob_in, w, "global_mlp.", L, act, ln), d_dyn.view(E, B, N, -1)


# ---- normalisers ------------------------------------------------------------------------------------------------------------
def update_normalizer_state(model, name, data):
    """Normalizer_v2.update (torch_helpers.py:774-781) or Normalizer.update (:821-836) on data [rows, dim]."""
    data = np.asarray(data, np.float64 if model.normalize_w_running_stats else np.float32)
    if data.shape[-1] == 0:
        return
    eps = model.epsilon
    if model.normalize_w_running_stats:
        ex = model._norm_extra.setdefault(name, {})
        if "sum" not in ex:
            ex.update(sum=np.zeros(data.shape[-1]), sum_of_squares=np.zeros(data.shape[-1]), count=1.0)  # re_init(): count starts at 1
        ex["sum"] = ex["sum"] + data.sum(0)
        ex["sum_of_squares"] = ex["sum_of_squares"] + np.square(data).sum(0)
        ex["count"] = ex["count"] + data.shape[0]
        mean = ex["sum"] / ex["count"]
        std = np.maximum(eps, np.sqrt(ex["sum_of_squares"] / ex["count"] - np.square(ex["sum"] / ex["count"]) + eps))
    else:
        t = torch.from_numpy(np.ascontiguousarray(data, np.float32))
        mean, std = t.mean(0).numpy(), t.std(0).numpy()
        std = np.where(std < eps, 1.0, std)
    model._normalizers[name] = (mean.astype(np.float32), std.astype(np.float32))
    model.weights_version += 1


class _EnsembleIterator:
    """TorchTrainingIterator with ensemble=True (torch_helpers.py:344-445): index arrays [E, batch] per step."""

    def __init__(self, size, E):
        self.size, self.E = size, E

    def epochs_bootstrapped(self, batch_size, epochs):
        sub = np.random.choice(self.size, (self.E, self.size), replace=True)
        for _ in range(epochs):
            perm = np.asarray([np.random.permutation(x) for x in sub])
            for j in range(int(np.ceil(self.size / batch_size))):
                yield perm[:, j * batch_size:(j + 1) * batch_size]

    def epochs_non_bootstrapped(self, batch_size, epochs):
        for _ in range(epochs):
            sub = np.asarray([np.random.permutation(self.size) for _ in range(self.E)])
            for j in range(int(np.ceil(self.size / batch_size))):
                yield sub[:, j * batch_size:(j + 1) * batch_size]

    def iterations(self, batch_size, n):
        sub = np.random.choice(self.size, (self.E, batch_size * n), replace=True)
        perm = np.asarray([np.random.permutation(x) for x in sub])
        for j in range(n):
            yield perm[:, j * batch_size:(j + 1) * batch_size]


class EnsembleTrainer:
    """Owns the trainable copy of a drop-in model's weights (torch Parameters on the CUDA device) and its optimizer."""

    def __init__(self, model, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("ceeus_b200 trains on a CUDA device; there is no CPU fallback")
        self.m = model
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.params = {k: torch.nn.Parameter(v.to(self.device, torch.float32).clone()) for k, v in model.state_dict().items()}
        tp = model.train_params
        self.epochs, self.iterations = tp.get("epochs"), tp.get("iterations")
        self.batch_size, self.lr = tp["batch_size"], tp["learning_rate"]
        self.weight_decay = tp.get("weight_decay", 0.0)
        self.bootstrapped = tp.get("bootstrapped", model.default_bootstrapped)
        self.latest_only = tp.get("train_epochs_only_with_latest_data", False)
        self.save_best_val = tp.get("save_best_val", False)
        self.improvement_threshold = tp.get("improvement_threshold", 0.01)
        self.grad_norm = tp.get("grad_norm")
        if not self.latest_only:
            assert not (self.epochs and self.iterations), "epochs and iterations together need train_epochs_only_with_latest_data"
        opt = tp.get("optimizer", "Adam")
        kw = {"eps": 0.0001} if opt == "Adam" else {}  # gnn_ensemble.py:176-187, mlp_ensemble.py:157-167
        self.optimizer = getattr(torch.optim, opt)(list(self.params.values()), lr=self.lr, weight_decay=self.weight_decay, **kw)
        if model._ckpt_extra.get("optimizer") is not None:
            try:  # continue from a checkpoint written by the reference (parameter order = state_dict order there as here)
                self.optimizer.load_state_dict(model._ckpt_extra["optimizer"])
            except (ValueError, KeyError):
                pass

    # -- model-specific pieces ---------------------------------------------------------------------------------------------
    def _nz(self, name):
        m, s = self.m._normalizer_list_by_name()[name]
        return torch.from_numpy(m).to(self.device), torch.from_numpy(s).to(self.device)

    def _split(self, obs):
        env = self.m.env
        A, D, S, N = env.agent_dim, env.object_dyn_dim, env.object_stat_dim, env.nObj
        agent = obs[..., :A]
        dyn = obs[..., A:A + N * D].reshape(*obs.shape[:-1], N, D)
        stat = obs[..., A + N * D:A + N * (D + S)].reshape(*obs.shape[:-1], N, S)
        return agent, dyn, stat

    def loss_and_per_model(self, obs, act, next_obs):
        """obs / act / next_obs [E,B,*] (goal dims are cut off here: env.obs_preproc) -> per-member MSE [E]."""
        m = self.m
        sd = m.env.observation_space_size_preproc
        obs, next_obs = obs[..., :sd], next_obs[..., :sd]
        if m.model_kind == "mlp":  # mlp_ensemble.py:293-330, 212-221
            x = torch.cat([obs, act], dim=-1)
            mu, sg = self._nz("in")
            y = ensemble_mlp((x - mu) / sg, self.params, "", m.num_layers, m.act_fn, m.layer_norm)
            mo, so = self._nz("out")
            target = ((next_obs - obs) - mo) / so
            return F.mse_loss(y, target, reduction="none").mean((1, 2))
        agent, dyn, stat = self._split(obs)  # gnn_ensemble.py:344-426
        agent_n, dyn_n, _ = self._split(next_obs)
        (ma, sa), (md, sdd), (ms, ss), (mu, su) = self._nz("in_agent"), self._nz("in_dyn"), self._nz("in_stat"), self._nz("in_action")
        (moa, soa), (mod, sod) = self._nz("out_agent"), self._nz("out_dyn")
        d_agent, d_dyn = gnn_forward(self.params, m, (agent - ma) / sa, (dyn - md) / sdd, (stat - ms) / ss if stat.shape[-1] else stat,
                                     (act - mu) / su)
        t_agent, t_dyn = ((agent_n - agent) - moa) / soa, ((dyn_n - dyn) - mod) / sod
        E, B = obs.shape[:2]
        out = torch.cat([d_agent, d_dyn.reshape(E, B, -1)], dim=-1)
        tgt = torch.cat([t_agent, t_dyn.reshape(E, B, -1)], dim=-1)
        return F.mse_loss(out, tgt, reduction="none").mean((1, 2))

    # -- the loop (gnn_ensemble.py:639-720, mlp_ensemble.py:293-429) ------------------------------------------------------------
    def _fit(self, observations, actions, next_observations, eval_buffer, mode):
        dev = self.device
        obs = torch.as_tensor(np.asarray(observations), dtype=torch.float32, device=dev)
        act = torch.as_tensor(np.asarray(actions), dtype=torch.float32, device=dev)
        nxt = torch.as_tensor(np.asarray(next_observations), dtype=torch.float32, device=dev)
        E = self.m.ensemble_size
        it = _EnsembleIterator(obs.shape[0], E)
        if mode == "epochs":
            batches = (it.epochs_bootstrapped if self.bootstrapped else it.epochs_non_bootstrapped)(self.batch_size, self.epochs)
        else:
            batches = it.iterations(self.batch_size, self.iterations)
        ev = None
        if eval_buffer:
            ev = tuple(torch.as_tensor(np.asarray(eval_buffer[k]), dtype=torch.float32, device=dev).unsqueeze(0).expand(E, -1, -1)
                       for k in ("observations", "actions", "next_observations"))
        epoch_length = int(np.ceil(obs.shape[0] / self.batch_size))
        best_val, best_weights = None, None
        if ev is not None and self.m.model_kind == "gnn":
            with torch.no_grad():
                best_val = self.loss_and_per_model(*ev).sum()
        accum, epoch_train_loss, test_loss = 0.0, 0.0, 0.0
        for i, idx in enumerate(batches):
            idx_t = torch.as_tensor(idx, device=dev)
            self.optimizer.zero_grad()
            loss = self.loss_and_per_model(obs[idx_t], act[idx_t], nxt[idx_t]).sum()  # mean over batch / dims, SUM over members
            loss.backward()
            if self.grad_norm is not None:
                torch.nn.utils.clip_grad_norm_(list(self.params.values()), self.grad_norm)
            self.optimizer.step()
            accum += loss.item()
            if (i + 1) % epoch_length == 0 or i == 0:
                if ev is not None:
                    if self.m.model_kind == "mlp":
                        # the dense model evaluates through a shuffled iterator (mlp_ensemble.py:223-253): the loss is the mean over
                        # the whole set either way, but the E permutations it draws advance np.random between training batches
                        for _ in range(E):
                            np.random.permutation(ev[0].shape[1])
                    with torch.no_grad():
                        test_loss = self.loss_and_per_model(*ev).sum()
                    if self.save_best_val:  # abstract_models.py maybe_get_best_weights: relative improvement above the threshold
                        if best_val is None or float((best_val - test_loss) / torch.abs(best_val)) > self.improvement_threshold:
                            best_val = test_loss if best_val is None else torch.minimum(best_val, test_loss)
                            best_weights = {k: v.detach().clone() for k, v in self.params.items()}
                epoch_train_loss = accum / min(epoch_length, i + 1)
                accum = 0.0
        if self.save_best_val and best_weights is not None:
            with torch.no_grad():
                for k, v in best_weights.items():
                    self.params[k].copy_(v)
        return {"epoch_train_loss": epoch_train_loss, "epoch_test_loss": test_loss}

    def train(self, rollout_buffer, eval_buffer=None, maybe_update_normalizer=True):
        if not (self.epochs or self.iterations):
            return None
        m = self.m
        if maybe_update_normalizer:
            m.update_normalizer(rollout_buffer.latest_rollouts if m.normalize_w_running_stats else rollout_buffer)
        src = rollout_buffer.latest_rollouts if self.latest_only else rollout_buffer
        stats = self._fit(src["observations"], src["actions"], src["next_observations"], eval_buffer, "epochs") if self.epochs else {}
        if self.iterations:
            old = self._fit(rollout_buffer["observations"], rollout_buffer["actions"], rollout_buffer["next_observations"], eval_buffer,
                            "iterations")
            stats = {"train_" + k + "_latest_data": v for k, v in stats.items()}
            stats.update({"train_" + k + "_old_data": v for k, v in old.items()})
        m.load_state_dict({k: v.detach() for k, v in self.params.items()})  # planner engines re-upload on their next call
        m._ckpt_extra["optimizer"] = self.optimizer.state_dict()
        return stats
