"""GPU: the sampling / elite / refit kernels and size-independent properties of the whole path."""
import numpy as np
import pytest
import torch

import icem_oracle as orc
from cases import CASES
from helpers import SeededNoise, build_inputs, product_controller, product_env, product_model, rel_err, step_noise

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _engine(P=64, H=30, n_obj=2, **kw):
    import ceeus_b200

    c = dict(env="construction", n_obj=n_obj, model="gnn", hidden=128, layers=2, wseed=3, identity_norm=True, oseed=2)
    spec, weights, norms, obs, goal = build_inputs(c)
    env = product_env(c, goal)
    model = product_model(c, env, weights, norms)
    args = dict(horizon=H, num_traj=P, sampler=dict(noise_beta=3.5), seed=1234)
    args.update(kw)
    eng = ceeus_b200.Engine(**args, **model.engine_kwargs())
    model.push_to(eng)
    eng.set_goal(goal)
    return eng, spec, obs, goal


@pytest.mark.parametrize("H", [20, 30, 35])
def test_device_sampler_matches_philox_oracle(H):
    """Philox4x32-10 counters are bit-exact; Box-Muller + irDFT agree to float tolerance."""
    P = 96
    eng, spec, _, _ = _engine(P=P, H=H)
    U, F = spec.action_dim, H // 2 + 1
    mean = torch.zeros(1, H, U, device=DEV)
    std = torch.ones(1, H, U, device=DEV)
    gauss = torch.zeros(P, U, 2 * F, device=DEV)
    first_row = 1000
    acts = eng.sample(mean, std, P, gauss=gauss, write_gauss=True, stream_id=7, first_row=first_row)
    z = orc.philox_normals(1234, 7, (first_row + P) * U, 2 * F)[first_row * U:].reshape(P, U, 2 * F)
    np.testing.assert_allclose(gauss.cpu().numpy(), z, rtol=2e-5, atol=2e-6)
    y = (torch.from_numpy(z) @ torch.from_numpy(orc.irdft_matrix(3.5, H))).numpy()  # [P,U,H]
    want = np.clip(y.transpose(0, 2, 1), -1, 1)
    np.testing.assert_allclose(acts.cpu().numpy(), want, rtol=1e-4, atol=2e-5)
    # same transform as the reference's irfft formulation
    y2 = orc.colored_noise_from_gaussians(torch.from_numpy(z[..., :F].copy()), torch.from_numpy(z[..., F:].copy()), 3.5, H)
    np.testing.assert_allclose(y, y2.numpy(), rtol=1e-4, atol=2e-5)


def test_device_sampler_statistics_and_determinism():
    P, H = 8192, 30
    eng, spec, _, _ = _engine(P=P, H=H)
    U = spec.action_dim
    mean = torch.zeros(1, H, U, device=DEV)
    std = torch.full((1, H, U), 1e-3, device=DEV)  # tiny std: no clipping, y = actions / std
    a1 = eng.sample(mean, std, P, stream_id=3)
    a2 = eng.sample(mean, std, P, stream_id=3)
    a3 = eng.sample(mean, std, P, stream_id=4)
    assert torch.equal(a1, a2) and not torch.equal(a1, a3)
    y = (a1 / 1e-3).cpu().numpy()  # [P,H,U]
    # the reference's sigma (colored_noise.py:185-187) leaves out the DC term, whose scale is f[0]:=f[1] (:177-182), so
    # the "unit variance" series really has std sqrt(mean_t sum_k M[k,t]^2) (1.203 for beta=3.5, H=30); the device
    # sampler must reproduce exactly that
    M = orc.irdft_matrix(3.5, H).astype(np.float64)
    std_theory = float(np.sqrt((M ** 2).sum(0).mean()))
    assert 1.15 < std_theory < 1.25
    assert abs(y.mean()) < 0.02 and abs(y.std() - std_theory) < 0.03
    # power-law spectrum: periodogram slope ~ -beta over the interior frequencies
    spec_pow = (np.abs(np.fft.rfft(y, axis=1)) ** 2).mean(axis=(0, 2))[1:-1]
    f = np.fft.rfftfreq(H)[1:-1]
    slope = np.polyfit(np.log(f), np.log(spec_pow), 1)[0]
    assert abs(slope + 3.5) < 0.15
    # supplied noise passes through untouched, with the reference's op order and clipping
    noise = torch.randn(P, U, H, device=DEV)
    m2 = torch.rand(1, H, U, device=DEV) - 0.5
    s2 = torch.rand(1, H, U, device=DEV) + 0.5
    a4 = eng.sample(m2, s2, P, noise=noise)
    want = torch.clamp(noise.transpose(2, 1) * s2 + m2, -1, 1)
    assert torch.equal(a4, want)


def test_white_noise_branch():
    eng, spec, _, _ = _engine(P=4096, H=12, sampler=dict(noise_beta=3.5, colored_noise=False))
    U = spec.action_dim
    a = eng.sample(torch.zeros(1, 12, U, device=DEV), torch.full((1, 12, U), 1e-3, device=DEV), 4096) / 1e-3
    assert abs(float(a.mean())) < 0.02 and abs(float(a.std()) - 1.0) < 0.02
    lag1 = float((a[:, 1:] * a[:, :-1]).mean())
    assert abs(lag1) < 0.02


def test_full_size_properties_config2():
    """BASELINE config 2 sizes (P=4096, H=30, N=2): properties that need no oracle."""
    c = dict(kind="mpc", env="construction", n_obj=2, model="gnn", hidden=128, layers=2, P=4096, H=30, wseed=5,
             identity_norm=True, running=False, oseed=1, curiosity=True, cost="sum", sampler={})
    spec, weights, norms, obs, goal = build_inputs(c)
    env = product_env(c, goal)
    model = product_model(c, env, weights, norms)
    ctrl = product_controller(c, env, model, seed=11)
    ctrl.beginning_of_rollout(observation=obs, state=None, mode="train")
    a = ctrl.get_action(obs, None)
    e = ctrl.engine
    assert a.shape == (4,) and np.all(np.abs(a) <= 1.0) and np.all(np.isfinite(a))
    traj = e.materialise()  # the planner itself keeps no trajectories (costs are reduced inside the rollout kernel)
    assert torch.isfinite(traj).all()
    # traj[0] is the start observation for every member / trajectory; every predicted obs carries env.goal
    assert torch.equal(traj[0], torch.from_numpy(obs).to(DEV).expand_as(traj[0]))
    assert torch.equal(traj[1:, :, :, spec.state_dim:], torch.from_numpy(goal).to(DEV).expand_as(traj[1:, :, :, spec.state_dim:]))
    # elites: ascending costs; fresh-sample elites agree with costs_ at their index; executed action = best elite's first
    ec, ei = e.elite_costs_[0].cpu().numpy(), e.elite_idx_[0].cpu().numpy()
    assert np.all(np.diff(ec) >= 0)
    costs = ctrl.costs_.cpu().numpy()
    fresh = ei < 4096
    np.testing.assert_array_equal(ec[fresh], costs[ei[fresh]])
    assert ec[0] <= costs.min()
    np.testing.assert_array_equal(a, e.elite_actions_[0, 0, 0].cpu().numpy())
    # sharding invariance of the rollout: the two halves rolled out separately equal the full launch bit for bit
    acts = e.actions_.clone()
    start = torch.from_numpy(obs).view(1, -1).to(DEV)
    full = e.rollout(start, acts).clone()
    halves = [model._rollout_engine(2048, 30).rollout(start, acts[i * 2048:(i + 1) * 2048].contiguous()).clone() for i in (0, 1)]
    assert torch.equal(full, torch.cat(halves, dim=2))
    # permutation equivariance over trajectories
    perm = torch.randperm(4096, device=DEV)
    assert torch.equal(e.rollout(start, acts[perm].contiguous()), full[:, :, perm])
    # determinism of a whole MPC step under the device RNG
    ctrl2 = product_controller(c, env, model, seed=11)
    ctrl2.beginning_of_rollout(observation=obs, state=None, mode="train")
    np.testing.assert_array_equal(ctrl2.get_action(obs, None), a)
    ctrl3 = product_controller(c, env, model, seed=12)
    ctrl3.beginning_of_rollout(observation=obs, state=None, mode="train")
    assert not np.array_equal(ctrl3.get_action(obs, None), a)


@pytest.mark.parametrize("precision", ["fp32", "tc"])
def test_multi_env_equals_independent_controllers(precision):
    """num_envs = 3 batched planning == 3 independent single-env controllers (config-4 semantics), bit for
    bit in both precision modes: a trajectory's result does not depend on the tile row it lands in."""
    c = dict(CASES["mpc_gnn_pg4_curious"])
    c["P"], c["H"] = 32, 6
    spec, weights, norms, obs, goal = build_inputs(c)
    rs = np.random.RandomState(0)
    obss = np.stack([obs + np.concatenate([0.1 * rs.standard_normal(spec.state_dim), np.zeros(spec.goal_dim)]).astype(np.float32)
                     for _ in range(3)])
    goals = np.stack([goal + np.float32(0.1 * i) for i in range(3)])
    obss[:, spec.state_dim:] = goals
    env = product_env(c, goal)
    env.goal = goals
    model = product_model(c, env, weights, norms, precision=precision)
    batched = product_controller(c, env, model, precision=precision, num_envs=3)
    batched.beginning_of_rollout(observation=obss, state=None, mode="train")
    noise = SeededNoise(5, 3.5)
    P, U, H = c["P"], spec.action_dim, c["H"]
    singles = []
    for i in range(3):
        env_i = product_env(c, goals[i])
        ctl = product_controller(c, env_i, product_model(c, env_i, weights, norms, precision=precision), precision=precision)
        ctl.beginning_of_rollout(observation=obss[i], state=None, mode="train")
        singles.append(ctl)
    for step in range(2):
        per_env = [[noise((P, U, H)) for _ in range(3)] for _ in range(3)]  # [env][iter]
        flat_b = torch.cat([torch.cat([per_env[e][it] for e in range(3)]).reshape(-1) for it in range(3)]).to(DEV)
        a_b = batched.get_action(obss, None, noise=flat_b)
        for i in range(3):
            flat_i = torch.cat([per_env[i][it].reshape(-1) for it in range(3)]).to(DEV)
            a_i = singles[i].get_action(obss[i], None, noise=flat_i)
            np.testing.assert_array_equal(a_b[i], a_i)
            assert torch.equal(batched.engine.mean_[i], singles[i].engine.mean_[0])
            assert torch.equal(batched.engine.elite_idx_[i], singles[i].engine.elite_idx_[0])


@pytest.mark.parametrize("precision", ["fp32", "tc"])
@pytest.mark.parametrize("name", ["mpc_gnn_c2_curious", "mpc_gnn_c6_stack_extr"])
def test_sharded_ranks_equal_single_gpu(name, precision):
    """world_size = 2 emulated on one GPU: two rank-engines run their local stage one after the other (no waiting on
    each other), their candidate records are concatenated as the all-gather would, both run the merge.  Result must
    equal the single-GPU planner bit for bit, and both ranks must agree."""
    import ceeus_b200

    c = CASES[name]
    spec, weights, norms, obs, goal = build_inputs(c)
    env = product_env(c, goal)
    model = product_model(c, env, weights, norms, precision=precision)
    single = product_controller(c, env, model, precision=precision)
    single.beginning_of_rollout(observation=obs, state=None, mode="train")
    sp = dict(single.engine.sampler)
    kw = dict(horizon=c["H"], num_traj=c["P"] // 2, sampler=sp, cost_along_trajectory=c["cost"],
              curiosity=c["curiosity"], extrinsic_reward=c.get("extrinsic", False), world_size=2)
    ranks = [ceeus_b200.Engine(rank=r, **kw, **model.engine_kwargs()) for r in (0, 1)]
    for e in ranks:
        model.push_to(e)
        e.set_goal(goal)
        e.begin_rollout()
        e.obs_dev_.copy_(torch.from_numpy(obs).view(1, -1))
    noise = SeededNoise(c["nseed"], 3.5)
    P, U, H, half = c["P"], spec.action_dim, c["H"], c["P"] // 2
    for step in range(c["steps"]):
        has = single.engine.has_elites()
        chunks = []
        for i in range(3):
            chunks.append(noise((P, U, H)))
            if i == 0 and sp["shift_elites_over_time"] and has and single.engine.n_kept > 0:
                chunks.append(noise((single.engine.n_kept, U, H)))
        flat = torch.cat([x.reshape(-1) for x in chunks]).to(DEV)
        a = single.get_action(obs, None, noise=flat)
        it = iter(chunks)
        for i in range(3):
            full = next(it).to(DEV)
            shift = next(it).to(DEV).contiguous() if (i == 0 and sp["shift_elites_over_time"] and has and single.engine.n_kept > 0) else None
            for r, e in enumerate(ranks):
                e.plan_iter_local(i, full[r * half:(r + 1) * half].contiguous(), shift)
            cand_all = torch.stack([e.cand_ for e in ranks]).contiguous()
            for e in ranks:
                e.plan_iter_update(i, cand_all)
        for e in ranks:
            e.plan_finish()
            assert torch.equal(e.action_out_[0].cpu(), torch.from_numpy(a))
            assert torch.equal(e.mean_, single.engine.mean_) and torch.equal(e.std_, single.engine.std_)
            assert torch.equal(e.elite_idx_, single.engine.elite_idx_)
            assert torch.equal(e.elite_actions_, single.engine.elite_actions_)
