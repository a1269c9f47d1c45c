"""Generates tests/golden/*.npz from the reference's OWN unmodified code (run in the build container only).

    python tests/golden/make_golden.py            # rewrites the fixtures
    python tests/golden/make_golden.py --check    # regenerates in memory and compares with the committed files

The reference (``/root/reference/mbrl``) is imported behind the ``oracle/ref_shims.py`` stubs; models, the
iCEM controllers, the environments' ``cost_fn`` and the colored-noise generator are the reference's classes.
Inputs are seeded through ``oracle/icem_oracle.py``'s ``synth_*`` helpers (numpy ``RandomState`` -- frozen
legacy stream, identical on every machine), loaded into the reference through ``load_state_dict`` and the
normalisers' own attributes.  Noise enters at the ``torch_powerlaw_psd_gaussian`` boundary.

Fixtures hold only small arrays (inputs that are not derivable from a seed, and the reference's outputs).
"""
import argparse
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

sys.path.insert(0, HERE)
import icem_oracle as orc  # noqa: E402
import ref_harness as rh  # noqa: E402
from cases import CASES, SeededNoise, build_inputs, cost_trajectories, rollout_actions  # noqa: E402


def build_reference(c, spec, weights, norms, goal):
    if c["env"] == "construction":
        env = rh.make_construction_env(c["n_obj"], case=c.get("case", "Singletower"), sparse=c.get("sparse", False),
                                       shaped_reward=c.get("shaped", True), goal=goal)
    else:
        env = rh.make_playground_env(c["n_obj"], goal=goal)
    if "abounds" in c:
        from gym import spaces

        env.action_space = spaces.Box(np.asarray(c["abounds"][0], np.float32), np.asarray(c["abounds"][1], np.float32),
                                      dtype="float32")
    if c["model"] == "gnn":
        model = rh.make_gnn_model(env, hidden_dim=c["hidden"], num_layers=c["layers"], running_stats=c["running"],
                                  ignore_agent_object=c.get("variant") != "edge_global",
                                  agent_as_global_node=c.get("variant") != "homog", embedding_dim=c.get("emb"),
                                  num_message_passing=c.get("n_mp", 1))
        rh.load_into_reference(model, weights, norms, c["running"])
    else:
        model = rh.make_mlp_model(env, hidden_dim=c["hidden"], num_layers=c["layers"], running_stats=c["running"])
        rh.load_into_reference(model, weights, norms, c["running"])
    return env, model


def gen_rollout_case(c):
    rh._prepare()
    from mbrl.controllers.abstract_controller import OpenLoopPolicy
    from mbrl.controllers.mpc_torch_curiosity import TorchCuriosityMpcICem

    spec, weights, norms, obs, goal = build_inputs(c)
    env, model = build_reference(c, spec, weights, norms, goal)
    acts = rollout_actions(c, spec)
    model.num_simulated_trajectories, model.horizon = c["P"], c["H"]
    model.preallocate_memory()
    start = torch.from_numpy(obs).view(1, -1).expand(c["P"], -1).contiguous()
    buf, _ = model.predict_n_steps(start_observations=start, start_states=[None] * c["P"],
                                   policy=OpenLoopPolicy(torch.from_numpy(acts)), horizon=c["H"])
    o, n, a = buf.as_array("observations"), buf.as_array("next_observations"), buf.as_array("actions")
    out = dict(next_observations=n.contiguous().numpy().copy(), first_observation=o[:, :, 0].contiguous().numpy().copy())
    assert torch.equal(o[:, :, 1:], n[:, :, :-1])
    assert torch.equal(a[:, 0], torch.from_numpy(acts))
    # disagreement bonus through the reference's own method
    ctrl = rh.make_controller(env, model, curiosity=True, horizon=c["H"], num_traj=c["P"])
    TorchCuriosityMpcICem._model_epistemic_costs(ctrl, buf)
    out["disagreement_bonus"] = ctrl._epistemic_bonus_per_path.numpy().copy()
    out["env_cost"] = env.cost_fn(o, a, n).numpy().copy()
    return out


def gen_cost_case(c):
    """env.cost_fn of the reference on the synthetic near-goal trajectories of cases.cost_trajectories."""
    rh._prepare()
    spec, weights, norms, obs, goal = build_inputs(c)
    if c["env"] == "robodesk":
        env = rh.make_robodesk_env(c["task"], c["reward"])
    else:
        env, _ = build_reference(c, spec, weights, norms, goal)
    traj = torch.from_numpy(cost_trajectories(c, spec, obs, goal))
    o, n = traj[:-1].permute(2, 1, 0, 3), traj[1:].permute(2, 1, 0, 3)
    a = torch.zeros(c["P"], 5, c["H"], spec.action_dim)
    return dict(env_cost=np.asarray(env.cost_fn(o, a, n).numpy(), np.float32).copy())


def gen_mpc_case(c):
    spec, weights, norms, obs, goal = build_inputs(c)
    env, model = build_reference(c, spec, weights, norms, goal)
    ctrl = rh.make_controller(env, model, curiosity=c["curiosity"], horizon=c["H"], num_traj=c["P"],
                              cost_along_trajectory=c["cost"], sampler=c["sampler"],
                              extrinsic_reward=c.get("extrinsic", False))
    ctrl.fully_deterministic = True  # stable sort (mpc_torch.py:451-454) so elite order is well defined
    rec = []
    orig_update = ctrl.update_distributions

    def tapped(paths, costs):
        item = dict(costs=costs.clone().numpy(), costs_per_model=ctrl.costs_per_model_.clone().numpy())
        orig_update(paths, costs)
        item["mean"] = ctrl.mean_.clone().numpy()
        item["std"] = ctrl.std_.clone().numpy()
        item["elite_actions"] = ctrl.elite_samples.as_array("actions")[:, 0].clone().numpy()
        item["elite_idx"] = torch.sort(costs, stable=True)[1][: ctrl.num_elites].numpy().astype(np.int64)
        all_act = paths.as_array("actions")[:, 0]
        assert torch.equal(all_act[torch.from_numpy(item["elite_idx"])], torch.from_numpy(item["elite_actions"]))
        rec.append(item)

    ctrl.update_distributions = tapped
    noise = SeededNoise(c["nseed"], 3.5)
    out = {}
    o = obs.copy()
    with rh.NoiseTap(lambda beta, size: noise(size)):
        ctrl.beginning_of_rollout(observation=o, state=None, mode="train")
        for s in range(c["steps"]):
            rec.clear()
            a = ctrl.get_action(o, None)
            out[f"s{s}_action"] = np.array(a, dtype=np.float32, copy=True)  # the reference returns a view of the model buffer
            out[f"s{s}_mean_after"] = ctrl.mean_.clone().numpy()
            for i, item in enumerate(rec):
                for k, v in item.items():
                    out[f"s{s}_i{i}_{k}"] = v
            # synthetic "environment": the next real observation is the best member-0 prediction, perturbed
            rs = np.random.RandomState(c["oseed"] + 50 + s)
            o = o.copy()
            o[: spec.state_dim] += (0.01 * rs.standard_normal(spec.state_dim)).astype(np.float32)
            out[f"s{s + 1}_obs"] = o.copy()
    out["s0_obs"] = obs.copy()
    return out


def gen_noise_case():
    """colored_noise.torch_powerlaw_psd_gaussian with torch.normal replaced by z*std (z supplied)."""
    rh._prepare()
    from mbrl.controllers import colored_noise

    out = {}
    for H in (20, 30, 35):
        F = H // 2 + 1
        rs = np.random.RandomState(900 + H)
        zr, zi = rs.standard_normal((5, 3, F)).astype(np.float32), rs.standard_normal((5, 3, F)).astype(np.float32)
        feed = [torch.from_numpy(zr), torch.from_numpy(zi)]
        orig = torch.normal
        torch.normal = lambda mean, std, *a, **k: feed.pop(0) * std
        try:
            y = colored_noise.torch_powerlaw_psd_gaussian(3.5, (5, 3, H))
        finally:
            torch.normal = orig
        out[f"H{H}_zr"], out[f"H{H}_zi"], out[f"H{H}_y"] = zr, zi, y.numpy().copy()
    return out


def generate_all(only=None):
    res = {}
    for name, c in CASES.items():
        if only and name not in only:
            continue
        res[name] = {"rollout": gen_rollout_case, "mpc": gen_mpc_case, "cost": gen_cost_case}[c["kind"]](c)
    if not only or "colored_noise" in only:
        res["colored_noise"] = gen_noise_case()
    return res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--only", nargs="*", default=None, help="fixture names to (re)generate; default: all")
    args = ap.parse_args()
    torch.set_num_threads(8)
    data = generate_all(args.only)
    for name, arrays in data.items():
        path = os.path.join(HERE, name + ".npz")
        if args.check:
            old = np.load(path)
            for k, v in arrays.items():
                np.testing.assert_allclose(old[k], v, rtol=1e-5, atol=1e-6, err_msg=f"{name}:{k}")
            print("ok", name)
        else:
            np.savez_compressed(path, **arrays)
            print("wrote", path, os.path.getsize(path))
