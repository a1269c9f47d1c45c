"""GPU: the drop-ins built from the reference's own settings files and fed from a live reference model object / a checkpoint
written by the reference, compared with the reference's own predictions (the reference runs on the CPU of the GPU box from
oracle/_ref/reference_mbrl.zip)."""
import functools
import os
import sys
import tempfile

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_harness as rh  # noqa: E402
from helpers import rel_err  # noqa: E402
from test_cpu_boundary import load_settings  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not rh.reference_available(), reason="reference package not available")]


@pytest.mark.parametrize("rel,env_kind,n_obj", [
    ("construction/curious_exploration/gnn_ensemble_cee_us.yaml", "construction", 2),
    ("playground/curious_exploration/gnn_ensemble_cee_us.yaml", "playground", 4),
    ("construction/curious_exploration/mlp_ensemble_cee_us.yaml", "construction", 4)])
def test_dropins_from_shipped_settings_track_a_live_reference_model(rel, env_kind, n_obj, monkeypatch):
    import ceeus_b200
    from ceeus_b200 import registry

    monkeypatch.setattr(torch, "load", functools.partial(torch.load, weights_only=False))
    params = load_settings(rel)
    fm_params, ctrl_params = dict(params["forward_model_params"]), dict(params["controller_params"])
    ref_env = rh.make_construction_env(n_obj) if env_kind == "construction" else rh.make_playground_env(n_obj)
    env = ceeus_b200.construction_env(n_obj) if env_kind == "construction" else ceeus_b200.playground_env(n_obj)
    # the reference's own model built from the same yaml dict, random init, normalisers updated once ("trained" state)
    rh._prepare()
    from mbrl.models import forward_model_from_string

    torch.manual_seed(0)
    ref = forward_model_from_string(params["forward_model"])(env=ref_env, **fm_params)
    rs = np.random.RandomState(1)
    O, U = ref_env.observation_space.shape[0], ref_env.action_space.shape[0]
    sd = ref_env.observation_space_size_preproc
    is_gnn = hasattr(ref, "input_normalizer_agent")
    if is_gnn:
        for nz in (ref.input_normalizer_agent, ref.input_normalizer_obj_dyn, ref.input_normalizer_obj_stat, ref.action_normalizer,
                   ref.output_agent_normalizer, ref.output_obj_normalizer):
            if nz.shape:
                nz.update((rs.standard_normal((64, nz.shape)) * 0.5 + 0.2).astype(np.float32))
    else:
        ref.input_normalizer.update((rs.standard_normal((64, sd + U)) * 0.5 + 0.2).astype(np.float32))
        ref.output_normalizer.update((rs.standard_normal((64, sd)) * 0.1).astype(np.float32))
    # drop-in from the registry key, same yaml dict
    key = params["forward_model"] + "B200"
    model = registry.forward_model_from_string(key)(env=env, **fm_params)
    model.sync_from(ref)
    E, B = model.ensemble_size, 7
    goal = (rs.standard_normal(O - sd) * 0.1).astype(np.float32)
    ref_env.goal = goal.copy()
    env.goal = goal.copy()
    obs = (rs.standard_normal((E, B, O)) * 0.3).astype(np.float32)
    obs[..., sd:] = goal
    act = rs.uniform(-1, 1, (E, B, U)).astype(np.float32)
    want, _, _ = ref.predict(observations=torch.from_numpy(obs), states=None, actions=torch.from_numpy(act))
    got, _, _ = model.predict(observations=torch.from_numpy(obs), states=None, actions=torch.from_numpy(act))
    assert rel_err(got.cpu().numpy(), want.numpy()) < 1e-5
    # the same through a checkpoint written by the reference's save()
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "ckpt.pt")
        ref.save(path)
        model2 = registry.forward_model_from_string(key)(env=env, **fm_params)
        model2.load(path)
        got2, _, _ = model2.predict(observations=torch.from_numpy(obs), states=None, actions=torch.from_numpy(act))
        assert torch.equal(got2, got)
    # ... and the controller from the yaml `controller` key / `controller_params`: a planning step runs
    ckey = {"mpc-curiosity-icem-torch": "mpc-curiosity-icem-b200", "mpc-icem-torch": "mpc-icem-b200"}[params["controller"]]
    ctrl = registry.controller_from_string(ckey)(env=env, forward_model=model, **dict(ctrl_params, logging=False))
    start = obs[0, 0]
    ctrl.beginning_of_rollout(observation=start, state=None, mode="train")
    a = ctrl.get_action(start, None)
    assert a.shape == (U,) and np.all(np.isfinite(a)) and np.all(np.abs(a) <= 1.0)
    assert ctrl.num_sim_traj == ctrl_params["num_simulated_trajectories"] and ctrl.horizon == ctrl_params["horizon"]
    assert tuple(ctrl.mean_.shape) == (ctrl_params["horizon"], U)
    es = ctrl.elite_samples
    assert tuple(es["actions"].shape) == (ctrl.num_elites, E, ctrl_params["horizon"], U)
    nxt = es["next_observations"]
    assert tuple(nxt.shape) == (ctrl.num_elites, E, ctrl_params["horizon"], O)
    # elite_samples re-simulates from the REAL observation (the library keeps its own device copy): first obs == start
    first = es["observations"][:, :, 0].cpu().numpy()
    np.testing.assert_array_equal(first, np.broadcast_to(start, first.shape))
