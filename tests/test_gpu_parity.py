"""GPU (B200): the CUDA path, called through the drop-in classes -> C ABI, against
  (1) the fixtures the REFERENCE produced (tests/golden), and
  (2) the CPU oracle on fresh seeded inputs,
with the fp32-mode bar of BASELINE.json: states / costs within 1e-5 relative, elite indices exact."""
import os

import numpy as np
import pytest
import torch

import icem_oracle as orc
from cases import CASES
from cases import cost_trajectories
from test_oracle_golden import expected_action
from helpers import (GOLD, SeededNoise, build_inputs, oracle_model, oracle_params, product_controller, product_env,
                     product_model, rel_err, rollout_actions, step_noise)

pytestmark = pytest.mark.gpu
TOL_FP32 = 1e-5  # BASELINE.json north_star: "1e-5 in fp32 mode" (relative, denominators floored at 1e-3)
# Costs computed from the product's OWN rollout inherit the state error amplified by the conditioning of the disagreement:
# std_e(x) is a difference of numbers of magnitude |x| ~ 1..30, so a state error of delta * |x| (delta <= 1e-5, the state
# bar) moves one std term by up to delta * |x| / std_e(x) relative -- 1e-4 .. 1e-3 for the near-deterministic dims; the sum
# over H x O terms averages most of it out (worst case measured over the fixtures: see DESIGN.md section 6).  The bar for the cost
# ARITHMETIC itself is 1e-5 and is checked on identical inputs (the fixture's trajectories) in every test below.
TOL_COST_OWN_ROLLOUT = 5e-5
DEV = "cuda:0"


class _Policy:  # OpenLoopPolicy stand-in (abstract_controller.py:178-220): carries action_sequences [P,H,U]
    def __init__(self, a):
        self.action_sequences = a


@pytest.mark.parametrize("name", [k for k, c in CASES.items() if c["kind"] == "rollout"])
def test_rollout_matches_reference_fixture(name):
    c = CASES[name]
    g = np.load(os.path.join(GOLD, name + ".npz"))
    spec, weights, norms, obs, goal = build_inputs(c)
    env = product_env(c, goal)
    model = product_model(c, env, weights, norms)
    acts = torch.from_numpy(rollout_actions(c, spec)).to(DEV)
    start = torch.from_numpy(obs).view(1, -1).expand(c["P"], -1).contiguous().to(DEV)
    buf, _ = model.predict_n_steps(start_observations=start, start_states=[None] * c["P"], policy=_Policy(acts),
                                   horizon=c["H"])
    nxt = buf.as_array("next_observations").cpu().numpy()
    assert nxt.shape == g["next_observations"].shape
    assert rel_err(nxt, g["next_observations"]) < TOL_FP32
    np.testing.assert_array_equal(buf.as_array("observations")[:, :, 0].cpu().numpy(), g["first_observation"])
    assert tuple(buf.as_array("actions").shape) == (c["P"], 5, c["H"], spec.action_dim)
    # costs on the device from the same trajectories
    eng = model._rollout_engine(c["P"], c["H"])
    traj = eng.traj_
    bonus = g["disagreement_bonus"]  # [P,H]
    cpm, costs = eng.costs(traj)  # engine default: curiosity, "sum"
    want = -bonus.sum(-1)
    assert rel_err(cpm.cpu().numpy(), np.repeat(want[:, None], 5, 1)) < TOL_COST_OWN_ROLLOUT
    assert rel_err(costs.cpu().numpy(), want) < TOL_COST_OWN_ROLLOUT
    # the cost kernel on the REFERENCE's trajectories (identical inputs): the 1e-5 bar proper
    ref_traj = torch.from_numpy(np.concatenate([g["first_observation"][:, :, None], g["next_observations"]], axis=2))
    cpm_r, costs_r = eng.costs(ref_traj.permute(2, 1, 0, 3).contiguous().to(DEV))
    assert rel_err(costs_r.cpu().numpy(), want) < TOL_FP32
    assert rel_err(cpm_r.cpu().numpy(), np.repeat(want[:, None], 5, 1)) < TOL_FP32


@pytest.mark.parametrize("name", [k for k, c in CASES.items() if c["kind"] == "rollout"])
@pytest.mark.parametrize("mode", ["task_sum", "task_best", "extr_best"])
def test_task_costs_match_reference_fixture(name, mode):
    import ceeus_b200

    c = CASES[name]
    g = np.load(os.path.join(GOLD, name + ".npz"))
    spec, weights, norms, obs, goal = build_inputs(c)
    env = product_env(c, goal)
    model = product_model(c, env, weights, norms)
    eng = ceeus_b200.Engine(horizon=c["H"], num_traj=c["P"], curiosity=mode.startswith("extr"),
                            extrinsic_reward=mode.startswith("extr"), extrinsic_reward_scale=0.7,
                            cost_along_trajectory="best" if mode.endswith("best") else "sum", **model.engine_kwargs())
    model.push_to(eng)
    eng.set_goal(goal)
    acts = torch.from_numpy(rollout_actions(c, spec)).to(DEV)
    start = torch.from_numpy(obs).view(1, -1).to(DEV)
    traj = eng.rollout(start, acts)
    cpm, costs = eng.costs(traj)
    envc = g["env_cost"].astype(np.float32)  # [P,E,H]
    path = envc if mode.startswith("task") else (-g["disagreement_bonus"][:, None, :] + np.float32(0.7) * envc)
    want = path[..., 1:].min(-1) if mode.endswith("best") else path.sum(-1)
    assert rel_err(cpm.cpu().numpy(), want) < TOL_COST_OWN_ROLLOUT
    assert rel_err(costs.cpu().numpy(), want.mean(-1)) < TOL_COST_OWN_ROLLOUT
    ref_traj = torch.from_numpy(np.concatenate([g["first_observation"][:, :, None], g["next_observations"]], axis=2))
    cpm_r, costs_r = eng.costs(ref_traj.permute(2, 1, 0, 3).contiguous().to(DEV))
    assert rel_err(cpm_r.cpu().numpy(), want) < TOL_FP32
    assert rel_err(costs_r.cpu().numpy(), want.mean(-1)) < TOL_FP32


@pytest.mark.parametrize("name", [k for k, c in CASES.items() if c["kind"] == "cost"])
@pytest.mark.parametrize("along", ["sum", "best"])
def test_task_cost_branches_on_near_goal_trajectories(name, along):
    """ceeus_costs on caller-supplied trajectories close to the goal: solved / unsolved blocks, the gripper mask, the sparse
    tower branch (fpp_construction_env.py:475-488) -- against env.cost_fn of the reference (fixture)."""
    import ceeus_b200

    c = CASES[name]
    g = np.load(os.path.join(GOLD, name + ".npz"))
    spec, weights, norms, obs, goal = build_inputs(c)
    env = product_env(c, goal)
    model = product_model(c, env, weights, norms)
    eng = ceeus_b200.Engine(horizon=c["H"], num_traj=c["P"], curiosity=False, cost_along_trajectory=along,
                            **model.engine_kwargs())
    eng.set_goal(goal)
    traj = torch.from_numpy(cost_trajectories(c, spec, obs, goal)).to(DEV)
    cpm, costs = eng.costs(traj)
    envc = g["env_cost"].astype(np.float32)
    want = envc[..., 1:].min(-1) if along == "best" else envc.sum(-1)
    assert rel_err(cpm.cpu().numpy(), want) < TOL_FP32
    assert rel_err(costs.cpu().numpy(), want.mean(-1)) < TOL_FP32


@pytest.mark.parametrize("name", [k for k, c in CASES.items() if c["kind"] == "mpc"])
def test_mpc_steps_match_reference_fixture(name):
    c = CASES[name]
    g = np.load(os.path.join(GOLD, name + ".npz"))
    spec, weights, norms, obs, goal = build_inputs(c)
    env = product_env(c, goal)
    model = product_model(c, env, weights, norms)
    ctrl = product_controller(c, env, model)
    noise = SeededNoise(c["nseed"], 3.5)
    ctrl.beginning_of_rollout(observation=g["s0_obs"], state=None, mode="train")
    e = ctrl.engine
    n_iter = ctrl.opt_iter
    for s in range(c["steps"]):
        flat, _ = step_noise(noise, c, spec, e.has_elites(), e.n_kept, DEV)
        a = ctrl.get_action(g[f"s{s}_obs"], None, noise=flat)
        i = n_iter - 1  # state after the last iteration is observable from outside
        np.testing.assert_array_equal(e.elite_idx_[0].cpu().numpy(), g[f"s{s}_i{i}_elite_idx"])
        k = e.k
        want_costs = np.sort(g[f"s{s}_i{i}_costs"], kind="stable")[:k]
        assert rel_err(e.elite_costs_[0].cpu().numpy(), want_costs) < TOL_COST_OWN_ROLLOUT
        assert rel_err(e.elite_actions_[0].cpu().numpy(), g[f"s{s}_i{i}_elite_actions"]) < TOL_FP32
        P = c["P"]
        assert rel_err(ctrl.costs_.cpu().numpy(), g[f"s{s}_i{i}_costs"][:P]) < TOL_COST_OWN_ROLLOUT
        np.testing.assert_allclose(a, expected_action(c, g, s), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(ctrl.mean_.cpu().numpy(), g[f"s{s}_mean_after"], rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("name", ["mpc_gnn_c2_task", "mpc_gnn_c6_stack_extr"])
def test_mpc_steps_with_env_cost_fn_called_on_materialised_rollouts(name):
    """An environment without a built-in device cost (SURVEY.md 8b "Env hand-off"): the controller materialises the rollout
    views and calls env.cost_fn itself (mpc_torch.py:121-125); everything else stays on the device.  Same fixtures as the
    built-in path: the result must not depend on where the task cost was evaluated."""
    c = CASES[name]
    g = np.load(os.path.join(GOLD, name + ".npz"))
    spec, weights, norms, obs, goal = build_inputs(c)
    env = product_env(c, goal)
    calls = []

    def cost_fn(observations, actions, next_observations):
        assert observations.is_cuda and tuple(observations.shape) == (c["P"], 5, c["H"], spec.obs_dim)
        assert tuple(actions.shape) == (c["P"], 5, c["H"], spec.action_dim)
        calls.append(1)
        return orc.construction_cost(spec, observations, next_observations)  # torch ops on the CUDA views

    env.cost_fn, env.force_external_cost = cost_fn, True
    model = product_model(c, env, weights, norms)
    ctrl = product_controller(c, env, model)
    assert ctrl.engine.external_cost
    noise = SeededNoise(c["nseed"], 3.5)
    ctrl.beginning_of_rollout(observation=g["s0_obs"], state=None, mode="train")
    e = ctrl.engine
    for s in range(c["steps"]):
        flat, _ = step_noise(noise, c, spec, e.has_elites(), e.n_kept, DEV)
        a = ctrl.get_action(g[f"s{s}_obs"], None, noise=flat)
        i = ctrl.opt_iter - 1
        np.testing.assert_array_equal(e.elite_idx_[0].cpu().numpy(), g[f"s{s}_i{i}_elite_idx"])
        assert rel_err(ctrl.costs_.cpu().numpy(), g[f"s{s}_i{i}_costs"][:c["P"]]) < TOL_COST_OWN_ROLLOUT
        np.testing.assert_allclose(a, g[f"s{s}_action"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(ctrl.mean_.cpu().numpy(), g[f"s{s}_mean_after"], rtol=1e-5, atol=2e-6)
    assert len(calls) == c["steps"] * ctrl.opt_iter


def _where(got, want):
    """Failure message: where the two trajectories [H+1,E,P,O] differ most."""
    d = np.abs(got - want)
    worst = np.unravel_index(np.argmax(d), d.shape)
    bad = np.argwhere(d > 1e-4 * (np.abs(want) + 1e-3))
    return (f"max |diff| {d.max():.3e} at (h,e,p,col)={worst}; {len(bad)} elements off by > 1e-4 relative; "
            f"h {sorted(set(bad[:, 0].tolist()))[:8]} e {sorted(set(bad[:, 1].tolist()))} "
            f"p {sorted(set(bad[:, 2].tolist()))[:16]} col {sorted(set(bad[:, 3].tolist()))[:16]}")


@pytest.mark.parametrize("env_kind,n_obj,model,hidden,P,H", [
    ("construction", 2, "gnn", 128, 300, 12), ("construction", 4, "gnn", 128, 130, 7),
    ("construction", 6, "gnn", 128, 77, 9), ("construction", 3, "gnn", 64, 64, 5),
    ("playground", 4, "gnn", 64, 200, 10), ("playground", 2, "gnn", 128, 33, 4),
    ("construction", 4, "mlp", 256, 257, 8), ("construction", 2, "mlp", 128, 70, 6)])
def test_rollout_matches_oracle_fresh_inputs(env_kind, n_obj, model, hidden, P, H):
    """Ragged sizes (P not a multiple of any tile), every supported N / hidden width, non-identity normalisers."""
    c = dict(kind="rollout", env=env_kind, n_obj=n_obj, model=model, hidden=hidden, layers=2 if model == "gnn" else 3,
             P=P, H=H, wseed=1000 + n_obj + hidden, identity_norm=False, running=False, oseed=77 + P)
    spec, weights, norms, obs, goal = build_inputs(c)
    env = product_env(c, goal)
    pm = product_model(c, env, weights, norms)
    om = oracle_model(c, spec, weights, norms)
    acts = rollout_actions(c, spec)
    rs = np.random.RandomState(5)
    starts = np.tile(obs, (P, 1))
    starts[:, :spec.state_dim] += (0.05 * rs.standard_normal((P, spec.state_dim))).astype(np.float32)
    want = orc.rollout(om, torch.from_numpy(starts), torch.from_numpy(acts), torch.from_numpy(goal)).numpy()
    buf, _ = pm.predict_n_steps(start_observations=torch.from_numpy(starts).to(DEV), start_states=None,
                                policy=_Policy(torch.from_numpy(acts).to(DEV)), horizon=H)
    got = pm._rollout_engine(P, H).traj_.cpu().numpy()
    assert got.shape == want.shape
    assert rel_err(got, want) < TOL_FP32, _where(got, want)


def test_predict_single_step_matches_oracle():
    c = CASES["rollout_gnn_pg4"]
    spec, weights, norms, obs, goal = build_inputs(c)
    env = product_env(c, goal)
    pm, om = product_model(c, env, weights, norms), oracle_model(c, spec, weights, norms)
    rs = np.random.RandomState(0)
    E, B = 5, 6
    o = np.tile(obs, (E, B, 1)) + (0.05 * rs.standard_normal((E, B, spec.obs_dim))).astype(np.float32)
    o[..., spec.state_dim:] = goal
    a = rs.uniform(-1, 1, (E, B, spec.action_dim)).astype(np.float32)
    nxt, _, _ = pm.predict(observations=torch.from_numpy(o), states=None, actions=torch.from_numpy(a))
    want = om.predict(torch.from_numpy(o[..., :spec.state_dim]), torch.from_numpy(a)).numpy()
    assert rel_err(nxt[..., :spec.state_dim].cpu().numpy(), want) < TOL_FP32
    np.testing.assert_array_equal(nxt[..., spec.state_dim:].cpu().numpy(), np.broadcast_to(goal, (E, B, goal.size)))


def test_mpc_matches_oracle_larger_population():
    """P = 1024, H = 30 (config-2 model, a quarter of its population): 2 MPC steps against the oracle."""
    c = dict(kind="mpc", env="construction", n_obj=2, model="gnn", hidden=128, layers=2, P=1024, H=30, wseed=5,
             identity_norm=True, running=False, oseed=1, nseed=9, steps=2, curiosity=True, cost="sum", sampler={})
    spec, weights, norms, obs, goal = build_inputs(c)
    env = product_env(c, goal)
    ctrl = product_controller(c, env, product_model(c, env, weights, norms))
    noise = SeededNoise(c["nseed"], 3.5)
    feed = {}
    oc = orc.ICEMOracle(oracle_model(c, spec, weights, norms), oracle_params(c), lambda size: feed["replay"](size))
    oc.set_goal(goal)
    oc.beginning_of_rollout()
    ctrl.beginning_of_rollout(observation=obs, state=None, mode="train")
    e = ctrl.engine
    for s in range(c["steps"]):
        flat, feed["replay"] = step_noise(noise, c, spec, e.has_elites(), e.n_kept, DEV)
        a = ctrl.get_action(obs, None, noise=flat)
        want = oc.get_action(obs)
        it = oc.trace[-1]["iters"][-1]
        np.testing.assert_array_equal(e.elite_idx_[0].cpu().numpy(), it["elite_idx"].numpy())
        assert rel_err(ctrl.costs_.cpu().numpy(), it["costs"].numpy()[: c["P"]]) < TOL_COST_OWN_ROLLOUT
        np.testing.assert_allclose(a, want, rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(ctrl.mean_.cpu().numpy(), oc.mean_.numpy(), rtol=1e-5, atol=2e-6)
        np.testing.assert_allclose(ctrl.std_.cpu().numpy(), oc.std_.numpy(), rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("tail", [False, True])
@pytest.mark.parametrize("precision", ["fp32", "tc"])
@pytest.mark.parametrize("name,num_envs", [("mpc_gnn_c2_curious", 1), ("mpc_gnn_c6_stack_extr", 1), ("mpc_gnn_c2_task", 1),
                                           ("mpc_gnn_pg4_curious", 3), ("mpc_mlp_c4_curious", 1),
                                           ("mpc_gnn_c3_task_coststd", 2)])
def test_planner_costs_from_compact_scratch_equal_cost_kernel_on_materialised_trajectories(name, num_envs, precision, tail):
    """The planner never materialises trajectories: the rollout kernel writes only the predicted dims into the library's compact
    scratch and the costs are reduced from there -- by a one-warp-per-trajectory launch (default) or, with
    ``cost_in_rollout_tail``, in the tail of the rollout kernel itself (no cost launch at all).  Same arithmetic as ceeus_costs on
    the fully materialised trajectories of the same population, up to the summation order over the (step, feature) terms."""
    import ceeus_b200

    c = dict(CASES[name])
    spec, weights, norms, obs, goal = build_inputs(c)
    rs = np.random.RandomState(3)
    obss = np.stack([obs + np.concatenate([0.05 * rs.standard_normal(spec.state_dim), np.zeros(spec.goal_dim)]).astype(np.float32)
                     for _ in range(num_envs)])
    goals = np.stack([goal + np.float32(0.01 * i) for i in range(num_envs)])
    obss[:, spec.state_dim:] = goals
    env = product_env(c, goal)
    env.goal = goals if num_envs > 1 else goal
    model = product_model(c, env, weights, norms, precision=precision)
    ctrl = product_controller(c, env, model, precision=precision, num_envs=num_envs, seed=5, cost_in_rollout_tail=tail)
    ctrl.beginning_of_rollout(observation=obss if num_envs > 1 else obss[0], state=None, mode="train")
    e = ctrl.engine
    assert not e.external_cost and e._traj is None  # no full-trajectory buffer exists on the planning path
    e.obs_dev_.copy_(torch.from_numpy(obss))
    e.set_goal(env.goal)
    model.push_to(e)
    l0 = e.kernel_launches()
    e.plan_iter_local(0)  # sample (device RNG) + rollout with the fused cost tail + top-k
    launched = e.kernel_launches() - l0
    fused_cpm, fused_costs = e.costs_per_model_.clone(), e.costs_.clone()
    traj = e.materialise()  # same population, same start observations -> full [H+1,E,n,O] trajectories
    cpm, costs = e.costs(traj)
    assert torch.isfinite(costs).all()
    assert rel_err(fused_costs.reshape(-1).cpu().numpy(), costs.cpu().numpy()) < 2e-6
    assert rel_err(fused_cpm.reshape(-1, e.E).cpu().numpy(), cpm.cpu().numpy()) < 2e-6
    assert launched == (3 if tail else 4), launched  # sample, rollout (+ costs in its tail | + compact cost launch), top-k
    # arrival counters are self-resetting: a second launch over the same scratch gives the same answer
    e.plan_iter_local(0)
    assert torch.equal(e.costs_, fused_costs) and torch.equal(e.costs_per_model_, fused_cpm)  # and the tail is deterministic


@pytest.mark.parametrize("task,reward", [("open_slide", "dense"), ("push_green", "sparse"), ("flat_block_off_table", "dense")])
def test_robodesk_planning_step_on_device_matches_oracle(task, reward):
    """The RoboDesk zero-shot setting (robodesk/common/gnn_ensemble.yaml: GNN ensemble with hidden 256 over 8 objects of 13 + 6
    features, agent 24, action 8; controller_tasks.yaml: task cost) entirely on the device: hidden-256 fp32 rollout kernel, the
    task cost of robodesk_env.py:53-135 reduced on the device -- no env.cost_fn round trip.  One MPC step against the oracle."""
    c = dict(kind="mpc", env="robodesk", task=task, reward=reward, n_obj=8, model="gnn", hidden=256, layers=2, P=40, H=6,
             wseed=61, identity_norm=False, running=False, oseed=62, nseed=63, curiosity=False, cost="sum",
             sampler=dict(opt_iterations=2, init_std=0.5))
    spec, weights, norms, obs, goal = build_inputs(c)
    env = product_env(c, goal)
    ctrl = product_controller(c, env, product_model(c, env, weights, norms))
    assert not ctrl.engine.external_cost
    om = oracle_model(c, spec, weights, norms)
    feed = {}
    oc = orc.ICEMOracle(om, oracle_params(c), lambda size: feed["replay"](size))
    oc.set_goal(goal)
    oc.beginning_of_rollout()
    ctrl.beginning_of_rollout(observation=obs, state=None, mode="train")
    noise = SeededNoise(c["nseed"], 3.5)
    e = ctrl.engine
    for s in range(2):
        flat, feed["replay"] = step_noise(noise, c, spec, e.has_elites(), e.n_kept, DEV)
        a = ctrl.get_action(obs, None, noise=flat)
        want = oc.get_action(obs)
        it = oc.trace[-1]["iters"][-1]
        assert rel_err(ctrl.costs_.cpu().numpy(), it["costs"].numpy()[: c["P"]]) < TOL_COST_OWN_ROLLOUT
        np.testing.assert_array_equal(e.elite_idx_[0].cpu().numpy(), it["elite_idx"].numpy())
        np.testing.assert_allclose(a, want, rtol=1e-5, atol=1e-6)
