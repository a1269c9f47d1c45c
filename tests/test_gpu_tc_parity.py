"""GPU: the tensor-core rollout path (CEEUS_PREC_TC: fp16 operands = 10-bit mantissa like TF32, fp32 accumulate) against
the reference fixtures and the CPU oracle, at the tf32/bf16-mode bar of BASELINE.json (1e-3 relative):
  * predicted states: rel-L2 < 1e-3 (the survey's TF32 probe measured 5e-4 the same way), and no single element off by
    more than 5e-3 of its feature's magnitude;
  * costs: max relative error < 1e-3;
  * elite indices: exact wherever the reference costs do not tie within that tolerance."""
import os

import numpy as np
import pytest
import torch

import icem_oracle as orc
from cases import CASES
from helpers import (GOLD, SeededNoise, build_inputs, elites_consistent, oracle_model, oracle_params, product_controller,
                     product_env, product_model, rel_err, rel_l2, rollout_actions, step_noise)

pytestmark = [pytest.mark.gpu, pytest.mark.bulk_oracle]
DEV = "cuda:0"
TOL_TC = 1e-3


class _Policy:
    def __init__(self, a):
        self.action_sequences = a


# (the homogeneous GNN and ignore_agent_object=false run on the fp32 path only)
GNN_ROLLOUTS = [k for k, c in CASES.items() if c["kind"] == "rollout" and c["model"] == "gnn" and "variant" not in c]


@pytest.mark.parametrize("rows", ["node", "global"])
@pytest.mark.parametrize("name", GNN_ROLLOUTS)
def test_tc_rollout_matches_reference_fixture(name, rows):
    c = dict(CASES[name], tc_rows=rows)
    g = np.load(os.path.join(GOLD, name + ".npz"))
    spec, weights, norms, obs, goal = build_inputs(c)
    env = product_env(c, goal)
    model = product_model(c, env, weights, norms, precision="tc")
    acts = torch.from_numpy(rollout_actions(c, spec)).to(DEV)
    start = torch.from_numpy(obs).view(1, -1).expand(c["P"], -1).contiguous().to(DEV)
    buf, _ = model.predict_n_steps(start_observations=start, start_states=None, policy=_Policy(acts), horizon=c["H"])
    nxt = buf.as_array("next_observations").cpu().numpy()
    eng = model._rollout_engine(c["P"], c["H"])
    assert eng.tc_status() == 0
    assert rel_l2(nxt, g["next_observations"]) < TOL_TC
    assert rel_err(nxt, g["next_observations"]) < 5e-3
    np.testing.assert_array_equal(buf.as_array("observations")[:, :, 0].cpu().numpy(), g["first_observation"])
    cpm, costs = eng.costs(eng.traj_)
    want = -g["disagreement_bonus"].sum(-1)
    assert rel_err(costs.cpu().numpy(), want) < TOL_TC


@pytest.mark.parametrize("env_kind,n_obj,hidden,P,H", [
    ("construction", 2, 128, 300, 30), ("construction", 2, 128, 41, 3), ("construction", 4, 128, 130, 7),
    ("construction", 6, 128, 77, 30), ("construction", 3, 64, 64, 5), ("construction", 5, 128, 19, 4),
    ("playground", 4, 64, 200, 30), ("playground", 2, 128, 33, 4), ("construction", 2, 128, 3000, 10),
    ("construction", 8, 128, 50, 3), ("construction", 12, 128, 9, 2), ("playground", 7, 64, 70, 4),
    ("construction", 2, 128, 40, 1), ("construction", 6, 128, 5, 1), ("playground", 4, 64, 2, 1)])
@pytest.mark.parametrize("rows", ["node", "global"])
def test_tc_rollout_matches_oracle_fresh_inputs(env_kind, n_obj, hidden, P, H, rows):
    """Ragged populations (partial groups, empty lock-step groups), every N / hidden width, per-trajectory start states, on both
    tile row layouts."""
    c = dict(kind="rollout", env=env_kind, n_obj=n_obj, model="gnn", hidden=hidden, layers=2, P=P, H=H,
             wseed=2000 + n_obj + hidden, identity_norm=False, running=False, oseed=177 + P, tc_rows=rows)
    spec, weights, norms, obs, goal = build_inputs(c)
    env = product_env(c, goal)
    pm = product_model(c, env, weights, norms, precision="tc")
    om = oracle_model(c, spec, weights, norms)
    acts = rollout_actions(c, spec)
    rs = np.random.RandomState(5)
    starts = np.tile(obs, (P, 1))
    starts[:, :spec.state_dim] += (0.05 * rs.standard_normal((P, spec.state_dim))).astype(np.float32)
    want = orc.rollout(om, torch.from_numpy(starts), torch.from_numpy(acts), torch.from_numpy(goal)).numpy()
    pm.predict_n_steps(start_observations=torch.from_numpy(starts).to(DEV), start_states=None,
                       policy=_Policy(torch.from_numpy(acts).to(DEV)), horizon=H)
    eng = pm._rollout_engine(P, H)
    got = eng.traj_.cpu().numpy()
    assert eng.tc_status() == 0
    assert np.isfinite(got).all()
    np.testing.assert_array_equal(got[0], want[0])
    np.testing.assert_array_equal(got[..., spec.state_dim:], want[..., spec.state_dim:])
    assert rel_l2(got, want) < TOL_TC
    assert rel_err(got, want) < 5e-3


@pytest.mark.parametrize("name", ["mpc_gnn_c2_curious", "mpc_gnn_c6_stack_extr", "mpc_gnn_pg4_curious", "mpc_mlp_c4_curious"])
def test_tc_mpc_steps_against_oracle(name):
    """Whole MPC steps in TC mode.  The oracle is re-run from the product's own (mean, std, elites) at every step so
    that one legitimately ambiguous elite (cost tie within tolerance) cannot snowball into a different plan."""
    c = dict(CASES[name])
    c["sampler"] = dict(c["sampler"], opt_iterations=1)
    spec, weights, norms, obs, goal = build_inputs(c)
    env = product_env(c, goal)
    ctrl = product_controller(c, env, product_model(c, env, weights, norms, precision="tc"), precision="tc")
    om = oracle_model(c, spec, weights, norms)
    noise = SeededNoise(c["nseed"], 3.5)
    ctrl.beginning_of_rollout(observation=obs, state=None, mode="train")
    e = ctrl.engine
    for step in range(3):
        feed = {}
        oc = orc.ICEMOracle(om, oracle_params(c), lambda size: feed["replay"](size))
        oc.set_goal(goal)
        oc.beginning_of_rollout()
        oc.mean_, oc.std_ = e.mean_[0].cpu().clone(), e.std_[0].cpu().clone()
        if e.has_elites():
            oc.elite_actions = e.elite_actions_[0].cpu().clone()
        flat, feed["replay"] = step_noise(noise, c, spec, e.has_elites(), e.n_kept, DEV)
        a = ctrl.get_action(obs, None, noise=flat)
        assert e.tc_status() == 0
        oc.get_action(obs)
        it = oc.trace[-1]["iters"][0]
        ref_costs = it["costs"].numpy()
        assert rel_err(ctrl.costs_.cpu().numpy(), ref_costs[: c["P"]]) < TOL_TC
        idx = e.elite_idx_[0].cpu().numpy()
        assert elites_consistent(idx, ref_costs, e.k, 2 * TOL_TC), (idx, it["elite_idx"].numpy())
        if set(idx.tolist()) == set(it["elite_idx"].numpy().tolist()):
            np.testing.assert_allclose(ctrl.mean_.cpu().numpy(), oc.mean_.numpy(), rtol=1e-5, atol=2e-6)
        assert np.all(np.isfinite(a)) and np.all(np.abs(a) <= 1.0)


def test_tc_step_is_deterministic_and_unsupported_shapes_are_refused():
    """(The config-2 tensor-core step is compared with the fp32 path and the oracle in test_gpu_tc_tiling.py.)"""
    c = dict(kind="mpc", env="construction", n_obj=2, model="gnn", hidden=128, layers=2, P=4096, H=30, wseed=5,
             identity_norm=True, running=False, oseed=1, curiosity=True, cost="sum", sampler={})
    spec, weights, norms, obs, goal = build_inputs(c)
    env = product_env(c, goal)
    acts = []
    for _ in range(2):
        ctrl = product_controller(c, env, product_model(c, env, weights, norms, precision="tc"), precision="tc", seed=11)
        ctrl.beginning_of_rollout(observation=obs, state=None, mode="train")
        acts.append([ctrl.get_action(obs, None).copy() for _ in range(2)])
        assert ctrl.engine.tc_status() == 0
    np.testing.assert_array_equal(acts[0], acts[1])
    c5 = dict(CASES["rollout_mlp_c4"], hidden=64)  # the tensor-core MLP kernel needs hidden_dim 128 / 256
    spec5, w5, n5, o5, g5 = build_inputs(c5)
    with pytest.raises(RuntimeError, match="CEEUS_PREC_TC"):
        product_model(c5, product_env(c5, g5), w5, n5, precision="tc")._rollout_engine(8, 4)
    cv = CASES["rollout_gnn_c2_homog"]
    specv, wv, nv, ov, gv = build_inputs(cv)
    with pytest.raises(RuntimeError, match="FP32 path only"):
        product_model(cv, product_env(cv, gv), wv, nv, precision="tc")._rollout_engine(8, 4)
    with pytest.raises(ValueError, match="precision"):  # there is no tf32 alias: "tc" is fp16 operands / fp32 accumulate
        product_model(c, env, weights, norms, precision="tf32")._rollout_engine(8, 4)


@pytest.mark.parametrize("variant", ["signed_gamma", "zero_gamma"])
@pytest.mark.parametrize("env_kind,n_obj,hidden", [("construction", 2, 128), ("construction", 4, 128), ("playground", 4, 64)])
@pytest.mark.parametrize("rows", ["node", "global"])
def test_tc_rollout_with_trained_like_layernorm(env_kind, n_obj, hidden, variant, rows):
    """Non-trivial LayerNorm affine parameters.  `signed_gamma`: gamma of either sign, beta != 0 -> the pack-time folding
    (sign into this layer's weight columns, |gamma| into the consumer's rows, beta / |gamma| in the epilogue) is active;
    `zero_gamma`: one gamma is exactly 0, which cannot be folded -> the general epilogue path must be taken."""
    P, H = 97, 12
    c = dict(kind="rollout", env=env_kind, n_obj=n_obj, model="gnn", hidden=hidden, layers=2, P=P, H=H,
             wseed=4100 + n_obj + hidden, identity_norm=False, running=False, oseed=903, tc_rows=rows)
    spec, weights, norms, obs, goal = build_inputs(c)
    rs = np.random.RandomState(77)
    weights = dict(weights)
    for k in list(weights):
        if k.endswith(".gamma"):
            g = rs.uniform(0.5, 1.5, weights[k].shape) * rs.choice([-1.0, 1.0], weights[k].shape)
            if variant == "zero_gamma" and k.startswith("node_mlp.layers.0"):
                g[:, 3] = 0.0
            weights[k] = g.astype(np.float32)
        elif k.endswith(".beta"):
            weights[k] = (0.3 * rs.standard_normal(weights[k].shape)).astype(np.float32)
    env = product_env(c, goal)
    pm = product_model(c, env, weights, norms, precision="tc")
    om = oracle_model(c, spec, weights, norms)
    acts = rollout_actions(c, spec)
    starts = np.tile(obs, (P, 1))
    starts[:, :spec.state_dim] += (0.05 * rs.standard_normal((P, spec.state_dim))).astype(np.float32)
    want = orc.rollout(om, torch.from_numpy(starts), torch.from_numpy(acts), torch.from_numpy(goal)).numpy()
    pm.predict_n_steps(start_observations=torch.from_numpy(starts).to(DEV), start_states=None,
                       policy=_Policy(torch.from_numpy(acts).to(DEV)), horizon=H)
    eng = pm._rollout_engine(P, H)
    got = eng.traj_.cpu().numpy()
    assert eng.tc_status() == 0 and np.isfinite(got).all()
    assert rel_l2(got, want) < TOL_TC
    assert rel_err(got, want) < 5e-3


def test_tc_mlp_rollout_matches_reference_fixture():
    """Dense-MLP ensemble (mlp_ensemble_cee_us.yaml: 62-256-256-256-58, SiLU) on the tensor-core path."""
    name = "rollout_mlp_c4"
    c = CASES[name]
    g = np.load(os.path.join(GOLD, name + ".npz"))
    spec, weights, norms, obs, goal = build_inputs(c)
    env = product_env(c, goal)
    model = product_model(c, env, weights, norms, precision="tc")
    acts = torch.from_numpy(rollout_actions(c, spec)).to(DEV)
    start = torch.from_numpy(obs).view(1, -1).expand(c["P"], -1).contiguous().to(DEV)
    buf, _ = model.predict_n_steps(start_observations=start, start_states=None, policy=_Policy(acts), horizon=c["H"])
    nxt = buf.as_array("next_observations").cpu().numpy()
    eng = model._rollout_engine(c["P"], c["H"])
    assert eng.tc_status() == 0
    assert rel_l2(nxt, g["next_observations"]) < TOL_TC
    assert rel_err(nxt, g["next_observations"]) < 5e-3


@pytest.mark.parametrize("n_obj,hidden,layers,P,H", [(4, 256, 3, 300, 30), (2, 256, 3, 77, 5), (4, 128, 2, 1000, 8), (6, 256, 1, 130, 4)])
def test_tc_mlp_rollout_matches_oracle_fresh_inputs(n_obj, hidden, layers, P, H):
    c = dict(kind="rollout", env="construction", n_obj=n_obj, model="mlp", hidden=hidden, layers=layers, P=P, H=H,
             wseed=3000 + n_obj + hidden, identity_norm=False, running=False, oseed=55 + P)
    spec, weights, norms, obs, goal = build_inputs(c)
    env = product_env(c, goal)
    pm = product_model(c, env, weights, norms, precision="tc")
    om = oracle_model(c, spec, weights, norms)
    acts = rollout_actions(c, spec)
    rs = np.random.RandomState(6)
    starts = np.tile(obs, (P, 1))
    starts[:, :spec.state_dim] += (0.05 * rs.standard_normal((P, spec.state_dim))).astype(np.float32)
    want = orc.rollout(om, torch.from_numpy(starts), torch.from_numpy(acts), torch.from_numpy(goal)).numpy()
    pm.predict_n_steps(start_observations=torch.from_numpy(starts).to(DEV), start_states=None,
                       policy=_Policy(torch.from_numpy(acts).to(DEV)), horizon=H)
    eng = pm._rollout_engine(P, H)
    got = eng.traj_.cpu().numpy()
    assert eng.tc_status() == 0 and np.isfinite(got).all()
    np.testing.assert_array_equal(got[0], want[0])
    np.testing.assert_array_equal(got[..., spec.state_dim:], want[..., spec.state_dim:])
    assert rel_l2(got, want) < TOL_TC
    assert rel_err(got, want) < 5e-3


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-5), ("tc", TOL_TC)])
@pytest.mark.parametrize("act,layer_norm,aggr", [("relu", True, "sum"), ("silu", True, "mean"), ("tanh", False, "sum"),
                                                 ("relu", False, "mean")])
@pytest.mark.parametrize("rows", ["node", "global"])
def test_gnn_model_param_variants(act, layer_norm, aggr, precision, tol, rows):
    """The other values `model_params` can take on this path (gnn_modules.py:17-131: act_fn, layer_norm, aggr_fn): sum
    aggregation (the folded W3 then carries no 1/(N-1) scale), non-ReLU activations and no LayerNorm (general epilogue)."""
    import ceeus_b200

    n_obj, hidden, P, H = 3, 128, 70, 6
    c = dict(kind="rollout", env="construction", n_obj=n_obj, model="gnn", hidden=hidden, layers=2, P=P, H=H,
             wseed=5100, identity_norm=False, running=False, oseed=31)
    spec, weights, norms, obs, goal = build_inputs(c)
    if not layer_norm:
        weights = {k: v for k, v in weights.items() if not (k.endswith(".gamma") or k.endswith(".beta"))}
    env = product_env(c, goal)
    pm = ceeus_b200.B200GNNEnsemble(env=env, model_params=dict(n=5, hidden_dim=hidden, num_layers=2, layer_norm=layer_norm,
                                                                act_fn=act, aggr_fn=aggr),
                                    backend_params=dict(precision=precision, tc_row_layout=rows))
    pm.load_state_dict({k: torch.from_numpy(v) for k, v in weights.items()})
    pm.set_normalizers(norms)
    if precision == "fp32" and rows == "global":
        pytest.skip("the row layout only exists on the tensor-core path")
    om = orc.GNNEnsembleOracle(spec=spec, weights=weights, normalizers=norms, hidden=hidden, num_layers=2, act=act,
                               layer_norm=layer_norm, aggr=aggr)
    acts = rollout_actions(c, spec)
    starts = np.tile(obs, (P, 1))
    want = orc.rollout(om, torch.from_numpy(starts), torch.from_numpy(acts), torch.from_numpy(goal)).numpy()
    pm.predict_n_steps(start_observations=torch.from_numpy(starts).to(DEV), start_states=None,
                       policy=_Policy(torch.from_numpy(acts).to(DEV)), horizon=H)
    eng = pm._rollout_engine(P, H)
    got = eng.traj_.cpu().numpy()
    assert eng.tc_status() == 0 and np.isfinite(got).all()
    assert rel_l2(got, want) < tol
    assert rel_err(got, want) < 5 * tol


# ---- fp16 operand range: saturating conversions + CEEUS_TC_RANGE in the status word (include/ceeus_b200.h) ----------------
TC_RANGE = 0x40000000


@pytest.mark.parametrize("rows", ["node", "global"])
def test_tc_layernorm_free_model_full_horizon_stays_in_range(rows):
    """layer_norm=False at H=30 with observations of magnitude ~100: nothing bounds the hidden activations but the weights, the
    fp16 operands stay far inside their range -> no flag, and the rollout still meets the tensor-core bar."""
    import ceeus_b200

    n_obj, hidden, P, H = 2, 128, 200, 30
    c = dict(kind="rollout", env="construction", n_obj=n_obj, model="gnn", hidden=hidden, layers=2, P=P, H=H,
             wseed=5200, identity_norm=False, running=False, oseed=32)
    spec, weights, norms, obs, goal = build_inputs(c)
    weights = {k: v for k, v in weights.items() if not (k.endswith(".gamma") or k.endswith(".beta"))}
    env = product_env(c, goal)
    pm = ceeus_b200.B200GNNEnsemble(env=env, model_params=dict(n=5, hidden_dim=hidden, num_layers=2, layer_norm=False, act_fn="relu"),
                                    backend_params=dict(precision="tc", tc_row_layout=rows))
    pm.load_state_dict({k: torch.from_numpy(v) for k, v in weights.items()})
    pm.set_normalizers(norms)
    om = orc.GNNEnsembleOracle(spec=spec, weights=weights, normalizers=norms, hidden=hidden, num_layers=2, act="relu", layer_norm=False)
    acts = rollout_actions(c, spec)
    rs = np.random.RandomState(3)
    starts = np.tile(obs, (P, 1))
    starts[:, :spec.state_dim] += (100.0 * rs.standard_normal((P, spec.state_dim))).astype(np.float32)
    want = orc.rollout(om, torch.from_numpy(starts), torch.from_numpy(acts), torch.from_numpy(goal)).numpy()
    pm.predict_n_steps(start_observations=torch.from_numpy(starts).to(DEV), start_states=None,
                       policy=_Policy(torch.from_numpy(acts).to(DEV)), horizon=H)
    eng = pm._rollout_engine(P, H)
    got = eng.traj_.cpu().numpy()
    assert eng.tc_status() == 0 and np.isfinite(got).all()
    assert rel_l2(got, want) < TOL_TC


@pytest.mark.parametrize("model", ["gnn", "gnn-no-layernorm", "mlp"])
def test_tc_saturated_operands_raise_the_range_flag_and_stay_finite(model):
    """Observations far outside the fp16 range (1e7 after normalisation) / exploding LayerNorm-free activations: the conversions
    saturate (no inf / NaN anywhere in the rollout of a LayerNorm model or the dense MLP), the status word carries CEEUS_TC_RANGE, and a planning step refuses to return
    an action computed from saturated operands."""
    import ceeus_b200

    P, H = 96, 6
    gnn = model != "mlp"
    c = dict(kind="mpc", env="construction", n_obj=2 if gnn else 4, model="gnn" if gnn else "mlp", hidden=128 if gnn else 256,
             layers=2 if gnn else 3, P=P, H=H, wseed=5300, identity_norm=True, running=False, oseed=33, curiosity=True, cost="sum")
    spec, weights, norms, obs, goal = build_inputs(c)
    env = product_env(c, goal)
    if model == "gnn-no-layernorm":
        weights = {k: (v * 40.0 if k.endswith(".weight") or k.endswith(".w") else v) for k, v in weights.items()
                   if not (k.endswith(".gamma") or k.endswith(".beta"))}
        pm = ceeus_b200.B200GNNEnsemble(env=env, model_params=dict(n=5, hidden_dim=128, num_layers=2, layer_norm=False, act_fn="relu"),
                                        backend_params=dict(precision="tc"))
        pm.load_state_dict({k: torch.from_numpy(v) for k, v in weights.items()})
        pm.set_normalizers(norms)
        scale = 1.0e3
    else:
        pm = product_model(c, env, weights, norms, precision="tc")
        scale = 1.0e7
    acts = torch.from_numpy(rollout_actions(dict(c, kind="rollout"), spec)).to(DEV)
    start = np.tile(obs, (P, 1))
    start[:, :spec.state_dim] *= scale
    start[:, :spec.state_dim] += scale
    pm.predict_n_steps(start_observations=torch.from_numpy(start).to(DEV), start_states=None, policy=_Policy(acts), horizon=H)
    eng = pm._rollout_engine(P, H)
    st = eng.tc_status()
    assert st & TC_RANGE and (st & ~TC_RANGE) == 0, hex(st)
    if model != "gnn-no-layernorm":  # (there the fp16 SUMS of saturated activations may still overflow: flagged, not finite)
        assert torch.isfinite(eng.traj_).all()
    # the same through the planner: ceeus_plan_step fails loudly instead of returning an action
    ctrl = product_controller(c, env, pm, precision="tc", seed=3)
    big = (obs * scale + scale).astype(np.float32)
    ctrl.beginning_of_rollout(observation=big, state=None, mode="train")
    with pytest.raises(RuntimeError, match="saturated"):
        ctrl.get_action(big, None)
