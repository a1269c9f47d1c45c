"""GPU: the RND controller drop-in (cee-us_b200/rnd.py) against the reference's own RND module, RMS and train_rnd
(mbrl/controllers/mpc_torch_rnd.py), run on the CPU of the box."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_harness as rh  # noqa: E402
from helpers import build_inputs, product_env, product_model, rel_err  # noqa: E402
from test_gpu_training import _Buffer  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not rh.reference_available(), reason="reference package not available")]
MP = dict(num_layers_predictor=1, num_layers_target=1, rnd_hidden_dim=256, rnd_rep_dim=128, rnd_network_type="mlp")
TP = dict(learning_rate=1e-3, batch_size=64, rnd_epochs=True, rnd_num_epochs_or_its=2)
SAMPLER = dict(opt_iterations=2, elites_size=10, alpha=0.1, init_std=0.8, relative_init=True, execute_best_elite=True,
               keep_previous_elites=True, shift_elites_over_time=True, finetune_first_action=False, fraction_elites_reused=0.3,
               use_mean_actions=True, colored_noise=True, noise_beta=3.5, use_ensemble_cost_std=False)


def _reference_rnd(env):
    rh._prepare()
    from mbrl import torch_helpers
    from mbrl.controllers.mpc_torch_rnd import RMS, RND

    torch_helpers.device = torch.device("cpu")
    mp = dict(MP, agent_dim=env.agent_dim, object_dyn_dim=env.object_dyn_dim, object_stat_dim=env.object_stat_dim, nObj=env.nObj,
              obs_dim=env.observation_space_size_preproc)
    torch.manual_seed(3)
    return RND(mp), RMS(device=torch.device("cpu"))


@pytest.mark.parametrize("n_members", [1, 5])
def test_rnd_controller_matches_reference_rnd_module(n_members):
    import ceeus_b200

    c = dict(kind="mpc", env="construction", n_obj=2, model="gnn", hidden=128, layers=2, P=48, H=6, wseed=71, identity_norm=False,
             running=False, oseed=72)
    spec, weights, norms, obs, goal = build_inputs(c)
    weights = {k: v[:n_members] for k, v in weights.items()}
    env = product_env(c, goal)
    model = ceeus_b200.B200GNNEnsemble(env=env, model_params=dict(n=n_members, hidden_dim=128, num_layers=2))
    model.load_state_dict({k: torch.from_numpy(v) for k, v in weights.items()})
    model.set_normalizers(norms)
    ctrl = ceeus_b200.B200RndMpcICem(env=env, forward_model=model, horizon=c["H"], num_simulated_trajectories=c["P"],
                                     action_sampler_params=SAMPLER, cost_along_trajectory="sum", logging=False,
                                     model_params=dict(MP), train_params=dict(TP), backend_params=dict(seed=5))
    ref, ref_rms = _reference_rnd(env)
    # same networks / normaliser state on both sides
    ctrl.predictor.load_state_dict(ref.predictor.state_dict())
    ctrl.target.load_state_dict(ref.target.state_dict())
    rs = np.random.RandomState(0)
    sd = env.observation_space_size_preproc
    data = (rs.standard_normal((300, spec.obs_dim)) * 0.4 + 0.3).astype(np.float32)
    ref.obs_normalizer.update(data[:, :sd])
    ctrl.obs_normalizer.update(data[:, :sd])
    # 1. one planning iteration: costs per (trajectory, member) = - sum_h bonus, with the bonus of the REFERENCE's RND + RMS applied
    #    to the materialised rollout in the reference's order (all rows of the iteration at once)
    ctrl.beginning_of_rollout(observation=obs, state=None, mode="train")
    e = ctrl.engine
    e.obs_dev_.copy_(torch.from_numpy(obs).view(1, -1))
    e.set_goal(goal)
    model.push_to(e)
    e.plan_iter_rollout(0)
    view = ceeus_b200.RolloutView.from_device_buffers(e.traj_, e.actions_)
    cost = ctrl._cost_from_rollout(view)
    nxt = view["next_observations"].cpu()
    P, E, H, O = nxt.shape
    ref.eval()
    with torch.no_grad():
        err = ref(nxt.reshape(-1, O)[:, :sd])
        _, var = ref_rms(err)
        want = -(err / (torch.sqrt(var) + 1e-8)).view(P, E, H)
    assert rel_err(cost.cpu().numpy(), want.numpy()) < 2e-5
    e.plan_iter_costs(cost)
    assert rel_err(e.costs_per_model_[0].cpu().numpy(), want.sum(-1).numpy()) < 2e-5
    assert rel_err(e.costs_[0].cpu().numpy(), want.sum(-1).mean(-1).numpy()) < 2e-5
    # 2. a whole MPC step runs through get_action (two iterations, RMS carried along)
    a = ctrl.get_action(obs, None)
    assert a.shape == (4,) and np.all(np.isfinite(a)) and float(ctrl.intrinsic_reward_rms.n) > 2 * P * E * H
    # 3. train_rnd: same seeded batches as the reference's TrainingIterator -> same predictor
    buf = _Buffer(dict(observations=data, actions=np.zeros((300, 4), np.float32), next_observations=data))

    class _RefCtrl:  # the reference's train_rnd / update_rnd / _update_normalizer bound to its RND module (no forward model needed)
        pass

    from mbrl.controllers.mpc_torch_rnd import TorchRNDMpcICem

    rc = object.__new__(TorchRNDMpcICem)
    rc.rnd, rc.env = ref, rh.make_construction_env(2)
    rc.rnd_opt = torch.optim.Adam(ref.parameters(), lr=TP["learning_rate"])
    rc.batch_size, rc.epochs, rc.iterations = TP["batch_size"], TP["rnd_num_epochs_or_its"], 0
    rc.logger = type("L", (), {"log": staticmethod(lambda *a, **k: None)})()
    np.random.seed(9)
    want_stats = rc.train_rnd(buf)
    np.random.seed(9)
    got_stats = ctrl.train_rnd(buf)
    assert abs(got_stats["epoch_rnd_loss"] - want_stats["epoch_rnd_loss"]) < 1e-4 * max(1.0, abs(want_stats["epoch_rnd_loss"]))
    for k, v in ref.predictor.state_dict().items():
        np.testing.assert_allclose(ctrl.predictor.state_dict()[k].cpu().numpy(), v.numpy(), rtol=2e-3, atol=2e-5, err_msg=k)
