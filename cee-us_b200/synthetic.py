"""Synthetic workloads for benchmarks and smoke runs (no dataset, no checkpoint, no simulator): seeded random-init weights in
the reference's ``state_dict`` layout, identity or random normalisers, a start observation and a goal of the shapes
SURVEY.md 8d lists, and the algorithmic FLOP count of one model step.

Weight scale follows the reference initialiser N(0, (1/(2 sqrt(in)))^2) (mbrl/torch_helpers.py:637-675; the "truncation"
there is a no-op).  All streams are numpy ``RandomState`` (frozen legacy generator: identical on every machine).
"""
import math

import numpy as np


def gnn_weight_shapes(env, hidden, num_layers, E):
    """state_dict layout of GraphNeuralNetworkEnsemble (gnn_modules.py:78-115; MLPParallelEnsemble, simple.py:36-57)."""
    na, A, U, D = env.object_dyn_dim + env.object_stat_dim, env.agent_dim, env.action_space.shape[0], env.object_dyn_dim
    shapes = {}
    for name, din, dout in (("edge_mlp", 2 * na + A + U, hidden), ("node_mlp", na + hidden + U + A, D),
                            ("global_mlp", A + hidden + U, A)):
        dims = [din] + [hidden] * num_layers
        for l in range(num_layers):
            shapes[f"{name}.layers.{l}.0.weight"] = (E, dims[l], dims[l + 1])
            shapes[f"{name}.layers.{l}.0.bias"] = (E, dims[l + 1])
            shapes[f"{name}.layers.{l}.1.gamma"] = (E, dims[l + 1])
            shapes[f"{name}.layers.{l}.1.beta"] = (E, dims[l + 1])
        shapes[f"{name}.layers.{num_layers}.0.weight"] = (E, hidden, dout)
        shapes[f"{name}.layers.{num_layers}.0.bias"] = (E, dout)
    return shapes


def mlp_weight_shapes(env, hidden, num_layers, E):
    """state_dict layout of the dense model (mlp_ensemble.py:39-50)."""
    sd = env.observation_space_size_preproc
    dims = [sd + env.action_space.shape[0]] + [hidden] * num_layers + [sd]
    shapes = {}
    for l in range(num_layers + 1):
        shapes[f"layers.{l}.0.weight"] = (E, dims[l], dims[l + 1])
        shapes[f"layers.{l}.0.bias"] = (E, dims[l + 1])
    return shapes


def synth_weights(shapes, seed, plain_init=False):
    """Seeded weights; unless ``plain_init`` the biases / LayerNorm gamma / beta are randomised too (the reference's own
    initialiser leaves them at 0 / 1 / 0)."""
    rs = np.random.RandomState(seed)
    out = {}
    for k in sorted(shapes):
        shp = shapes[k]
        if k.endswith("weight"):
            out[k] = (rs.standard_normal(shp) / (2.0 * math.sqrt(shp[1]))).astype(np.float32)
        elif k.endswith("bias") or k.endswith("beta"):
            out[k] = np.zeros(shp, np.float32) if plain_init else (0.1 * rs.standard_normal(shp)).astype(np.float32)
        elif k.endswith("gamma"):
            out[k] = np.ones(shp, np.float32) if plain_init else (1.0 + 0.1 * rs.standard_normal(shp)).astype(np.float32)
        else:
            raise KeyError(k)
    return out


def normalizer_dims(env, model_kind):
    U = env.action_space.shape[0]
    if model_kind == "gnn":
        return {"in_agent": env.agent_dim, "in_dyn": env.object_dyn_dim, "in_stat": env.object_stat_dim, "in_action": U,
                "out_agent": env.agent_dim, "out_dyn": env.object_dyn_dim}
    sd = env.observation_space_size_preproc
    return {"in": sd + U, "out": sd}


def synth_normalizers(dims, seed, identity):
    """(mean, std) per group; identity == a model before its first train()."""
    rs = np.random.RandomState(seed + 7919)
    out = {}
    for k in sorted(dims):
        d = dims[k]
        if identity:
            out[k] = (np.zeros(d, np.float32), np.ones(d, np.float32))
        else:
            out[k] = ((0.1 * rs.standard_normal(d)).astype(np.float32), rs.uniform(0.5, 1.5, d).astype(np.float32))
    return out


def synth_start_obs(env, seed):
    """(observation [O], goal [G]) -- SURVEY.md 8d "Synthetic inputs"."""
    rs = np.random.RandomState(seed)
    N = env.nObj
    if env.kind == "construction":
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
    sd = env.observation_space_size_preproc
    assert obs.shape[0] == env.observation_space.shape[0]
    return obs, obs[sd:].copy()


def flops_per_model_step(env, model_kind, hidden, num_layers):
    """Algorithmic GEMM FLOPs (MACs x 2; LayerNorm, activations, aggregation, normalisers, costs excluded) of one
    (trajectory, ensemble member, horizon step) -- SURVEY.md 8d."""
    A, U, D = env.agent_dim, env.action_space.shape[0], env.object_dyn_dim
    if model_kind == "mlp":
        sd = env.observation_space_size_preproc
        dims = [sd + U] + [hidden] * num_layers + [sd]
        return 2 * sum(a * b for a, b in zip(dims[:-1], dims[1:]))
    na, N = D + env.object_stat_dim, env.nObj

    def mlp(din, dout):
        dims = [din] + [hidden] * num_layers + [dout]
        return sum(a * b for a, b in zip(dims[:-1], dims[1:]))

    return 2 * (N * (N - 1) * mlp(2 * na + A + U, hidden) + N * mlp(na + A + U + hidden, D) + mlp(A + U + hidden, A))
