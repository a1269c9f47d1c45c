"""GPU: the tensor-core rollout kernels on the code paths the BENCHMARK runs -- every BASELINE.json size and every tiling
branch of the kernels (trajectory groups taken one after the other by a warpgroup = ``rounds`` >= 2, the uneven 3/4 split
that keeps the MMA-issuing warp empty in every round = ``uneven`` 1 or in the leading rounds only = ``uneven`` 2, both tile row layouts of the GNN kernel: "node" = N rows per trajectory with separate
node / global stages, "global" = N + 1 rows with the node and global MLP K-stacked into one stage each) -- against the CPU
oracle at the tf32/bf16-mode bar (1e-3 relative).

Which branch a case runs is asserted through the library itself (``ceeus_tc_tiling``), not through a copy of the formula,
so a change of the tiling cannot silently take these cases off the branches they are here for."""
import numpy as np
import pytest
import torch

import icem_oracle as orc
from helpers import (build_inputs, elites_consistent, oracle_model, product_controller, product_env, product_model, rel_err,
                     rel_l2, rollout_actions)

pytestmark = [pytest.mark.gpu, pytest.mark.bulk_oracle]
DEV = "cuda:0"
TOL_TC = 1e-3


def _engine_for(c, weights, norms, goal, num_envs=1):
    import ceeus_b200

    env = product_env(c, goal if num_envs == 1 else np.tile(goal, (num_envs, 1)))
    model = product_model(c, env, weights, norms, precision="tc")
    eng = ceeus_b200.Engine(horizon=c["H"], num_traj=c["P"] // num_envs, num_envs=num_envs, **model.engine_kwargs())
    model.push_to(eng)
    return eng


def _check_rollout(c, want_rounds=None, want_uneven=None, num_envs=1, want_rows=None):
    """One launch over c['P'] trajectories (num_envs start observations / goals) against the oracle."""
    spec, weights, norms, obs, goal = build_inputs(c)
    P, H = c["P"], c["H"]
    eng = _engine_for(c, weights, norms, goal, num_envs)
    rounds, uneven = eng.tc_tiling(P)
    if c["model"] == "gnn":
        assert eng.tc_row_layout() == (want_rows or c.get("tc_rows", "node")), eng.tc_row_layout()
    if want_rounds is not None:
        assert rounds == want_rounds, (rounds, want_rounds)
    if want_uneven is not None:
        assert uneven == want_uneven, (uneven, want_uneven)
    om = oracle_model(c, spec, weights, norms)
    acts = rollout_actions(c, spec)
    rs = np.random.RandomState(17)
    per = P // num_envs
    if num_envs == 1:
        # every trajectory starts from its own state: a row that lands in the wrong tile slot cannot go unnoticed
        starts = np.tile(obs, (P, 1))
        starts[:, :spec.state_dim] += (0.05 * rs.standard_normal((P, spec.state_dim))).astype(np.float32)
        goals = goal[None]
        eng.set_goal(goal)
        want = orc.rollout(om, torch.from_numpy(starts), torch.from_numpy(acts), torch.from_numpy(goal)).numpy()
        start_dev = torch.from_numpy(starts).to(DEV)
    else:
        # config-4 semantics: num_envs planners batched in one launch, one start observation and one goal per environment
        env_obs = np.tile(obs, (num_envs, 1))
        env_obs[:, :spec.state_dim] += (0.1 * rs.standard_normal((num_envs, spec.state_dim))).astype(np.float32)
        goals = np.stack([goal + np.float32(0.05 * i) for i in range(num_envs)])
        env_obs[:, spec.state_dim:] = goals
        eng.set_goal(goals)
        want = np.concatenate([
            orc.rollout(om, torch.from_numpy(env_obs[i:i + 1]).expand(per, -1), torch.from_numpy(acts[i * per:(i + 1) * per]),
                        torch.from_numpy(goals[i])).numpy() for i in range(num_envs)], axis=2)
        start_dev = torch.from_numpy(env_obs).to(DEV)
    got = eng.rollout(start_dev, torch.from_numpy(acts).to(DEV)).cpu().numpy()
    assert eng.tc_status() == 0
    assert got.shape == want.shape and np.isfinite(got).all()
    np.testing.assert_array_equal(got[0], want[0])
    np.testing.assert_array_equal(got[..., spec.state_dim:], want[..., spec.state_dim:])
    assert rel_l2(got, want) < TOL_TC, rel_l2(got, want)
    assert rel_err(got, want) < 5e-3
    # per-group error: a single mis-tiled group of trajectories would be invisible in the global rel-L2
    d = np.linalg.norm((got - want).reshape(H + 1, -1, P, got.shape[-1]).transpose(2, 0, 1, 3).reshape(P, -1), axis=1)
    r = np.linalg.norm(want.transpose(2, 0, 1, 3).reshape(P, -1), axis=1)
    assert float(np.max(d / r)) < 5 * TOL_TC, float(np.max(d / r))
    # the device costs on the same trajectories against the oracle's (curiosity, "sum")
    cpm, costs = eng.costs(eng.traj_)
    nxt = torch.from_numpy(want[1:]).permute(2, 1, 0, 3)
    want_cost = -orc.disagreement_bonus(nxt).sum(-1).numpy()
    assert rel_err(costs.cpu().numpy().reshape(-1), want_cost) < TOL_TC
    return eng


# ---- BASELINE.json sizes, full horizon --------------------------------------------------------------------------
def test_tc_rollout_config2_full_size():
    """configs[1]: construction N=2, 4096 trajectories, H=30 on the layout the library picks for it -- global rows (N + 1 = 3 rows
    per trajectory, 5 stages per model step), two rounds per slot, the 14-pair member with one pure-issuer and one even round."""
    c = dict(kind="rollout", env="construction", n_obj=2, model="gnn", hidden=128, layers=2, P=4096, H=30, wseed=5,
             identity_norm=True, running=False, oseed=1, tc_rows="auto")
    _check_rollout(c, want_rounds=[2] * 5, want_uneven=[1, 1, 1, 1, 2], want_rows="global")


def test_tc_rollout_config2_full_size_node_rows():
    """configs[1] forced onto the node-major layout: two rounds per slot at 55 % tile fill."""
    c = dict(kind="rollout", env="construction", n_obj=2, model="gnn", hidden=128, layers=2, P=4096, H=30, wseed=5,
             identity_norm=True, running=False, oseed=1, tc_rows="node")
    _check_rollout(c, want_rounds=[2] * 5, want_uneven=[0] * 5)


def test_tc_rollout_config1_size_picks_global_rows():
    """configs[0] (the population the reference's yamls ship: 128 trajectories, H=20) is latency bound -- one round at 3 % fill --
    so the library picks the layout with fewer stages per step."""
    c = dict(kind="rollout", env="construction", n_obj=2, model="gnn", hidden=128, layers=2, P=128, H=20, wseed=5,
             identity_norm=True, running=False, oseed=1, tc_rows="auto")
    _check_rollout(c, want_rounds=[1] * 5, want_uneven=[0] * 5, want_rows="global")


def test_tc_rollout_config3_full_size():
    """configs[2] per GPU: construction N=6 (30 edges), 2048 trajectories, H=30 -> two rounds, uneven split for the members
    that own 15 CTA pairs, one pure-issuer round + one even round for the member that owns 14."""
    c = dict(kind="rollout", env="construction", n_obj=6, model="gnn", hidden=128, layers=2, P=2048, H=30, wseed=5,
             identity_norm=True, running=False, oseed=1, tc_rows="auto")
    _check_rollout(c, want_rounds=[2] * 5, want_uneven=[1, 1, 1, 1, 2], want_rows="node")


def test_tc_rollout_config4_full_size():
    """configs[3] per GPU: playground N=4, Hd=64, 8 environments x 1024 trajectories, H=30 -> three rounds, uneven split
    for the 14-pair member; per-environment start observation and goal."""
    c = dict(kind="rollout", env="playground", n_obj=4, model="gnn", hidden=64, layers=2, P=8192, H=30, wseed=5,
             identity_norm=False, running=True, oseed=1, tc_rows="auto")
    _check_rollout(c, want_rounds=[3] * 5, want_uneven=[0, 0, 0, 0, 1], num_envs=8, want_rows="node")


def test_tc_rollout_config4_full_size_global_rows():
    """configs[3] per GPU on the global-row layout (Hd = 64, four slots, 5 rows per trajectory): four rounds for the 14-pair member."""
    c = dict(kind="rollout", env="playground", n_obj=4, model="gnn", hidden=64, layers=2, P=8192, H=30, wseed=5,
             identity_norm=False, running=True, oseed=1, tc_rows="global")
    _check_rollout(c, want_rounds=[3, 3, 3, 3, 4], want_uneven=[2, 2, 2, 2, 1], num_envs=8)


def test_tc_mlp_rollout_config5_full_size():
    """configs[4]: dense-MLP ensemble 62-256x3-58 SiLU, 8192 trajectories, H=30 -> three 128-trajectory tiles per CTA."""
    c = dict(kind="rollout", env="construction", n_obj=4, model="mlp", hidden=256, layers=3, P=8192, H=30, wseed=5,
             identity_norm=True, running=False, oseed=1)
    _check_rollout(c, want_rounds=[3] * 5)


# ---- every (rounds, uneven) combination at a short horizon ---------------------------------------------------------
@pytest.mark.parametrize("env_kind,n_obj,hidden,P,H,rounds,uneven", [
    ("construction", 2, 128, 2886, 3, [1] * 5, [1] * 5),
    ("construction", 2, 128, 3138, 2, [1] * 5, [1, 1, 1, 1, 0]),
    ("construction", 2, 128, 3586, 2, [1, 1, 1, 1, 2], [0] * 5),
    ("construction", 2, 128, 5763, 2, [2] * 5, [1] * 5),
    ("construction", 2, 128, 6800, 2, [2] * 5, [2, 2, 2, 2, 0]),
    ("construction", 2, 128, 7681, 2, [3] * 5, [0] * 5),
    ("construction", 2, 128, 10100, 2, [3] * 5, [2] * 5),
    ("construction", 6, 128, 905, 3, [1] * 5, [1] * 5),
    ("construction", 6, 128, 1206, 3, [2] * 5, [0] * 5),
    ("construction", 6, 128, 1801, 2, [2] * 5, [1] * 5),
    ("construction", 6, 128, 2150, 2, [2] * 5, [2, 2, 2, 2, 0]),
    ("construction", 6, 128, 2403, 2, [3] * 5, [0] * 5),
    ("construction", 6, 128, 2704, 2, [3] * 5, [1] * 5),
    ("construction", 6, 128, 3160, 2, [3] * 5, [2] * 5),
    ("construction", 3, 128, 3607, 2, [2] * 5, [1] * 5),
    ("playground", 4, 64, 2886, 3, [1] * 5, [1] * 5),
    ("playground", 4, 64, 5763, 2, [2] * 5, [1] * 5),
    ("playground", 4, 64, 6800, 2, [2] * 5, [2, 2, 2, 2, 0]),
    ("playground", 4, 64, 7681, 2, [3] * 5, [0] * 5),
    ("playground", 4, 64, 8647, 2, [3] * 5, [1] * 5)])
def test_tc_rollout_every_tiling_branch(env_kind, n_obj, hidden, P, H, rounds, uneven):
    c = dict(kind="rollout", env=env_kind, n_obj=n_obj, model="gnn", hidden=hidden, layers=2, P=P, H=H,
             wseed=2300 + n_obj + hidden, identity_norm=False, running=False, oseed=400 + P, tc_rows="node")
    _check_rollout(c, want_rounds=rounds, want_uneven=uneven)


@pytest.mark.parametrize("env_kind,n_obj,hidden,P,H,rounds,uneven", [
    ("construction", 2, 128, 333, 4, [1] * 5, [0] * 5),
    ("construction", 2, 128, 1850, 3, [1] * 5, [1] * 5),
    ("construction", 2, 128, 2300, 2, [1, 1, 1, 1, 2], [0] * 5),
    ("construction", 2, 128, 3000, 2, [2] * 5, [0] * 5),
    ("construction", 2, 128, 3700, 2, [2] * 5, [1] * 5),
    ("construction", 2, 128, 4300, 2, [2] * 5, [2, 2, 2, 2, 0]),
    ("construction", 2, 128, 4900, 2, [3] * 5, [0] * 5),
    ("construction", 2, 128, 6350, 2, [3] * 5, [2] * 5),
    ("construction", 3, 128, 1500, 3, [1] * 5, [1] * 5),
    ("construction", 3, 128, 2000, 2, [2] * 5, [0] * 5),
    ("construction", 3, 128, 3400, 2, [2] * 5, [2, 2, 2, 2, 0]),
    ("construction", 6, 128, 750, 3, [1] * 5, [1] * 5),
    ("construction", 6, 128, 1000, 2, [2] * 5, [0] * 5),
    ("construction", 6, 128, 1500, 2, [2] * 5, [1] * 5),
    ("construction", 6, 128, 1700, 2, [2] * 5, [2, 2, 2, 2, 0]),
    ("playground", 4, 64, 2200, 3, [1] * 5, [1] * 5),
    ("playground", 4, 64, 3000, 2, [2] * 5, [0] * 5),
    ("playground", 4, 64, 4400, 2, [2] * 5, [1] * 5),
    ("playground", 4, 64, 5100, 2, [2] * 5, [2, 2, 2, 2, 0]),
    ("playground", 4, 64, 5800, 2, [3] * 5, [0] * 5)])
def test_tc_rollout_every_tiling_branch_global_rows(env_kind, n_obj, hidden, P, H, rounds, uneven):
    """The same tiling branches on the global-row layout (N + 1 tile rows per trajectory, K-stacked node / global stages)."""
    c = dict(kind="rollout", env=env_kind, n_obj=n_obj, model="gnn", hidden=hidden, layers=2, P=P, H=H,
             wseed=2300 + n_obj + hidden, identity_norm=False, running=False, oseed=400 + P, tc_rows="global")
    _check_rollout(c, want_rounds=rounds, want_uneven=uneven)


@pytest.mark.parametrize("P,H,rounds", [(3000, 3, [1] * 5), (3841, 2, [2] * 5), (5000, 2, [2] * 5), (7700, 2, [3] * 5)])
def test_tc_mlp_rollout_every_round_count(P, H, rounds):
    c = dict(kind="rollout", env="construction", n_obj=4, model="mlp", hidden=256, layers=3, P=P, H=H, wseed=3304,
             identity_norm=False, running=False, oseed=500 + P)
    _check_rollout(c, want_rounds=rounds)


# ---- whole MPC step at config-2 size: tensor-core mode against fp32 mode (same device RNG stream) and the oracle ----
def test_tc_full_step_config2_agrees_with_fp32_and_oracle():
    c = dict(kind="mpc", env="construction", n_obj=2, model="gnn", hidden=128, layers=2, P=4096, H=30, wseed=5,
             identity_norm=True, running=False, oseed=1, curiosity=True, cost="sum", sampler=dict(opt_iterations=1))
    spec, weights, norms, obs, goal = build_inputs(c)
    env = product_env(c, goal)
    ctrl = product_controller(c, env, product_model(c, env, weights, norms, precision="tc"), precision="tc", seed=11)
    ctrl.beginning_of_rollout(observation=obs, state=None, mode="train")
    a_tc = ctrl.get_action(obs, None)
    e = ctrl.engine
    assert e.tc_status() == 0
    assert e.tc_row_layout() == "global"  # the layout the benchmark's config-2 step runs (picked by the library)
    ctrl32 = product_controller(c, env, product_model(c, env, weights, norms), seed=11)
    ctrl32.beginning_of_rollout(observation=obs, state=None, mode="train")
    ctrl32.get_action(obs, None)
    e32 = ctrl32.engine
    # one iCEM iteration from the same (mean, std) and the same Philox stream: the sampled populations are identical
    assert torch.equal(e.actions_, e32.actions_)
    costs_tc, costs_32 = ctrl.costs_.cpu().numpy(), ctrl32.costs_.cpu().numpy()
    assert np.isfinite(costs_tc).all()
    assert rel_err(costs_tc, costs_32) < TOL_TC, rel_err(costs_tc, costs_32)
    assert elites_consistent(e.elite_idx_[0].cpu().numpy(), costs_32, e.k, 2 * TOL_TC)
    # ... and the oracle on the very same population
    om = oracle_model(c, spec, weights, norms)
    acts = e.actions_.cpu()
    traj = orc.rollout(om, torch.from_numpy(obs).view(1, -1).expand(c["P"], -1), acts, torch.from_numpy(goal))
    want = -orc.disagreement_bonus(traj[1:].permute(2, 1, 0, 3)).sum(-1).numpy()
    assert rel_err(costs_tc, want) < TOL_TC
    assert rel_err(costs_32, want) < 1e-5
    assert elites_consistent(e.elite_idx_[0].cpu().numpy(), want, e.k, 2 * TOL_TC)
    assert elites_consistent(e32.elite_idx_[0].cpu().numpy(), want, e.k, 2e-5)
    assert np.all(np.isfinite(a_tc)) and np.all(np.abs(a_tc) <= 1.0)


@pytest.mark.parametrize("cfg", ["c2", "c3", "c2-global", "c3-global"])
def test_tc_sharding_and_permutation_invariance_full_size(cfg):
    """A trajectory's tensor-core result does not depend on the launch it is part of or the tile row it lands in: halves
    rolled out separately (one round each at c2) equal the full launch (two rounds) bit for bit; so does a permutation."""
    n_obj, P = (2, 4096) if cfg.startswith("c2") else (6, 2048)
    H = 8
    c = dict(kind="rollout", env="construction", n_obj=n_obj, model="gnn", hidden=128, layers=2, P=P, H=H, wseed=5,
             identity_norm=True, running=False, oseed=1, tc_rows="global" if cfg.endswith("global") else "node")
    spec, weights, norms, obs, goal = build_inputs(c)
    eng = _engine_for(c, weights, norms, goal)
    eng.set_goal(goal)
    half = _engine_for(dict(c, P=P // 2), weights, norms, goal)
    half.set_goal(goal)
    assert eng.tc_tiling(P)[0] != half.tc_tiling(P // 2)[0]  # the two launches really tile differently
    acts = torch.from_numpy(rollout_actions(c, spec)).to(DEV)
    start = torch.from_numpy(obs).view(1, -1).to(DEV)
    full = eng.rollout(start, acts).clone()
    halves = [half.rollout(start, acts[i * (P // 2):(i + 1) * (P // 2)].contiguous()).clone() for i in (0, 1)]
    assert torch.equal(full, torch.cat(halves, dim=2))
    perm = torch.randperm(P, device=DEV)
    assert torch.equal(eng.rollout(start, acts[perm].contiguous()), full[:, :, perm])
    assert eng.tc_status() == 0 and half.tc_status() == 0
