"""Builders shared by the GPU parity tests, smoke() and bench.py: the product-side objects (ceeus_b200 drop-ins)
and the oracle-side objects for one case dict of tests/golden/cases.py."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

import icem_oracle as orc  # noqa: E402
from cases import SeededNoise, build_inputs, rollout_actions  # noqa: E402,F401

GOLD = os.path.join(ROOT, "tests", "golden")


def product_env(c, goal):
    import ceeus_b200

    if c["env"] == "construction":
        env = ceeus_b200.construction_env(c["n_obj"], case=c.get("case", "Singletower"), sparse=c.get("sparse", False),
                                          shaped_reward=c.get("shaped", True))
    elif c["env"] == "robodesk":
        env = ceeus_b200.robodesk_env(c["task"], c["reward"])
    else:
        env = ceeus_b200.playground_env(c["n_obj"])
    env.goal = np.asarray(goal, np.float32).copy()
    if "abounds" in c:
        env.action_space.low = np.asarray(c["abounds"][0], np.float32)
        env.action_space.high = np.asarray(c["abounds"][1], np.float32)
    return env


def product_model(c, env, weights, norms, precision="fp32"):
    import ceeus_b200

    if c["model"] == "gnn":
        mp = dict(n=5, hidden_dim=c["hidden"], num_layers=c["layers"], layer_norm=True, act_fn="relu")
        kw = {}
        if c.get("variant") == "edge_global":
            mp["ignore_agent_object"] = False
        elif c.get("variant") == "homog":
            mp.update(embedding_dim=c["emb"], num_message_passing=c.get("n_mp", 1))
            kw["agent_as_global_node"] = False
        m = ceeus_b200.B200GNNEnsemble(env=env, model_params=mp, normalize_w_running_stats=c.get("running", False),
                                       backend_params=dict(precision=precision, tc_row_layout=c.get("tc_rows", "auto")), **kw)
    else:
        m = ceeus_b200.B200MLPEnsemble(env=env, model_params=dict(n=5, hidden_dim=c["hidden"], num_layers=c["layers"],
                                                                   act_fn="silu", layer_norm=False),
                                       backend_params=dict(precision=precision))
    m.load_state_dict({k: torch.from_numpy(v) for k, v in weights.items()})
    m.set_normalizers(norms)
    return m


def product_controller(c, env, model, precision="fp32", **backend):
    import ceeus_b200

    sp = dict(opt_iterations=3, elites_size=10, alpha=0.1, init_std=0.8, relative_init=True, execute_best_elite=True,
              keep_previous_elites=True, shift_elites_over_time=True, finetune_first_action=False,
              fraction_elites_reused=0.3, use_mean_actions=True, colored_noise=True, noise_beta=3.5,
              use_ensemble_cost_std=False)
    sp.update(c.get("sampler", {}))
    kw = dict(env=env, forward_model=model, horizon=c["H"], num_simulated_trajectories=c["P"],
              action_sampler_params=sp, cost_along_trajectory=c.get("cost", "sum"), logging=False,
              backend_params=dict(precision=precision, tc_row_layout=c.get("tc_rows", "auto"), **backend))
    if c.get("curiosity", True):
        return ceeus_b200.B200CuriosityMpcICem(extrinsic_reward=c.get("extrinsic", False), **kw)
    return ceeus_b200.B200MpcICem(**kw)


def oracle_model(c, spec, weights, norms):
    cls = orc.GNNEnsembleOracle if c["model"] == "gnn" else orc.MLPEnsembleOracle
    kw = {}
    if c.get("variant") == "edge_global":
        kw["edge_global"] = True
    elif c.get("variant") == "homog":
        kw.update(homogeneous=True, num_message_passing=c.get("n_mp", 1))
    return cls(spec=spec, weights=weights, normalizers=norms, hidden=c["hidden"], num_layers=c["layers"], **kw)


def oracle_params(c):
    kw = dict(horizon=c["H"], num_traj=c["P"], cost_along_trajectory=c.get("cost", "sum"),
              curiosity=c.get("curiosity", True), extrinsic_reward=c.get("extrinsic", False))
    kw.update(c.get("sampler", {}))
    return orc.ICEMParams(**kw)


def step_noise(noise_fn, c, spec, has_elites, n_kept, device):
    """One MPC step worth of noise in the reference's draw order -> (flat device tensor for the product, replay fn for
    the oracle)."""
    P, U, H = c["P"], spec.action_dim, c["H"]
    sp = dict(opt_iterations=3, shift_elites_over_time=True)
    sp.update(c.get("sampler", {}))
    chunks = []
    for i in range(sp["opt_iterations"]):
        chunks.append(noise_fn((P, U, H)))
        if i == 0 and sp["shift_elites_over_time"] and has_elites and n_kept > 0:
            chunks.append(noise_fn((n_kept, U, H)))
    flat = torch.cat([x.reshape(-1) for x in chunks]).to(device)
    it = iter(chunks)

    def replay(size):
        x = next(it)
        assert tuple(x.shape) == tuple(size), (x.shape, size)
        return x

    return flat, replay


def rel_err(a, b):
    """Relative error used by the parity bars (1e-5 fp32 mode, 1e-3 tensor-core mode): the worse of
      * rel-L2 of the whole tensor, and
      * max_i |a_i - b_i| / (|b_i| + s_i), s_i = largest magnitude the same feature takes anywhere in b
        (last axis = feature for tensors with >= 3 axes, the whole tensor otherwise).
    An element-wise ratio without s_i is meaningless for features that pass through zero (velocities, deltas):
    fp32 rounding of state + delta is relative to the state's magnitude, not the element's."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if b.size == 0:
        return 0.0
    if b.ndim >= 3:
        scale = np.max(np.abs(b).reshape(-1, b.shape[-1]), axis=0)
    else:
        scale = np.max(np.abs(b))
    scale = np.maximum(scale, 1e-30)
    elem = float(np.max(np.abs(a - b) / (np.abs(b) + scale)))
    l2 = float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
    return max(elem, l2)


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def elites_consistent(idx, ref_costs, k, rel_tol):
    """Tolerance-aware elite comparison: every reference candidate that beats the k-th best cost by more
    than the precision-mode tolerance must be selected, none that is worse by more than it may be."""
    ref_costs = np.asarray(ref_costs, np.float64)
    kth = np.sort(ref_costs)[k - 1]
    margin = rel_tol * max(abs(kth), float(np.max(np.abs(ref_costs))) * 0.1)
    sure_in = set(np.nonzero(ref_costs < kth - margin)[0].tolist())
    sure_out = set(np.nonzero(ref_costs > kth + margin)[0].tolist())
    sel = set(int(i) for i in idx)
    return sure_in <= sel and not (sel & sure_out)
