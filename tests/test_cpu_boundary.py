"""CPU: the drop-in boundary against the REFERENCE's own interfaces, settings files and checkpoints.
Needs the reference package (``/root/reference`` in the build container, ``oracle/_ref/reference_mbrl.zip`` elsewhere)."""
import inspect
import os
import sys
import tempfile

import numpy as np
import pytest
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_harness as rh  # noqa: E402

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason="reference package not available")
SETTINGS = "experiments/cee_us/settings"


def _abstract_methods(cls):
    return sorted(n for n in getattr(cls, "__abstractmethods__", ()))


def test_dropins_implement_every_abstract_method_of_the_reference_interfaces():
    """mbrl/base_types.py:40-118: Controller.get_action; ForwardModel.train / reset / got_actual_observation_and_env_state /
    rollout_generator / rollout_field_names / predict / predict_n_steps / save / load -- same names, and every parameter of
    the abstract signature is accepted by the drop-in under the same name."""
    rh._prepare()
    import ceeus_b200
    from mbrl import base_types

    for abstract, dropins in ((base_types.ForwardModel, (ceeus_b200.B200GNNEnsemble, ceeus_b200.B200MLPEnsemble)),
                              (base_types.Controller, (ceeus_b200.B200MpcICem, ceeus_b200.B200CuriosityMpcICem))):
        names = _abstract_methods(abstract)
        assert names
        for cls in dropins:
            for name in names:
                assert callable(getattr(cls, name, None)), f"{cls.__name__} lacks {name}"
                want = inspect.signature(getattr(abstract, name)).parameters
                got = inspect.signature(getattr(cls, name)).parameters
                has_var_kw = any(p.kind is inspect.Parameter.VAR_KEYWORD for p in got.values())
                has_var_pos = any(p.kind is inspect.Parameter.VAR_POSITIONAL for p in got.values())
                # keyword-only parameters must exist under the same name; positional ones may be renamed (the reference's own
                # subclasses do: GNNForwardEnsembleModel.reset(obs) vs ForwardModel.reset(observation)) but must be as many
                n_pos_got = sum(p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD) for p in got.values())
                n_pos_want = sum(p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD) for p in want.values())
                assert n_pos_got >= n_pos_want or has_var_pos, f"{cls.__name__}.{name}: too few positional parameters"
                for pname, par in want.items():
                    if par.kind is par.KEYWORD_ONLY:
                        assert pname in got or has_var_kw, f"{cls.__name__}.{name} does not accept '{pname}'"
    # the controller surface the rollout loop uses beyond get_action (abstract_controller.py:47-62, mpc_torch_curiosity.py:82-95)
    for cls in (ceeus_b200.B200MpcICem, ceeus_b200.B200CuriosityMpcICem):
        for name in ("beginning_of_rollout", "end_of_rollout", "reset_horizon", "preallocate_memory"):
            assert callable(getattr(cls, name, None)), name
        sig = inspect.signature(cls.beginning_of_rollout).parameters
        assert {"observation", "state", "mode"} <= set(sig)


def _load_settings(rel):
    """yaml + the `__import__` lists of smart_settings (later files override earlier ones, the file's own keys last)."""
    def merge(a, b):
        for k, v in b.items():
            if isinstance(v, dict) and isinstance(a.get(k), dict):
                merge(a[k], v)
            else:
                a[k] = v
        return a

    def load(path):
        d = yaml.safe_load(rh.read_reference_text(path)) or {}
        out = {}
        imports = d.pop("__import__", [])
        for imp in [imports] if isinstance(imports, str) else imports:
            merge(out, load(imp))
        return merge(out, d)

    return load(SETTINGS + "/" + rel)


load_settings = _load_settings


@pytest.mark.parametrize("rel,ctrl_cls,model_cls", [
    ("construction/curious_exploration/gnn_ensemble_cee_us.yaml", "B200CuriosityMpcICem", "B200GNNEnsemble"),
    ("construction/curious_exploration/mlp_ensemble_cee_us.yaml", "B200CuriosityMpcICem", "B200MLPEnsemble"),
    ("playground/curious_exploration/gnn_ensemble_cee_us.yaml", "B200CuriosityMpcICem", "B200GNNEnsemble"),
    ("construction/zero_shot_generalization/gnn_ensemble_cee_us_zero_shot_stack.yaml", "B200MpcICem", "B200GNNEnsemble")])
def test_constructors_accept_the_shipped_settings(rel, ctrl_cls, model_cls):
    """Every key of the merged yaml `forward_model_params` / `controller_params` binds to the drop-in constructors."""
    import ceeus_b200

    if not rh.reference_has(SETTINGS + "/" + rel):
        pytest.skip("settings file not in this checkout")
    params = _load_settings(rel)
    fm = dict(params["forward_model_params"])
    inspect.signature(getattr(ceeus_b200, model_cls).__init__).bind(None, env=None, **fm)
    mp = inspect.signature(getattr(ceeus_b200, model_cls)._parse_model_params)
    mp.bind(None, **fm["model_params"])
    cp = dict(params["controller_params"])
    inspect.signature(getattr(ceeus_b200, ctrl_cls).__init__).bind(None, env=None, forward_model=None, **cp)
    # the sampler keys are the ones the engine knows (mpc_torch.py:480-512)
    from ceeus_b200.engine import SAMPLER_DEFAULTS

    assert set(cp["action_sampler_params"]) <= set(SAMPLER_DEFAULTS), set(cp["action_sampler_params"]) - set(SAMPLER_DEFAULTS)


@pytest.mark.parametrize("running", [False, True])
def test_checkpoint_round_trip_with_the_reference(running, monkeypatch):
    """A checkpoint written by the reference's save() loads into the drop-in; the drop-in's save() loads back into a fresh
    reference model with identical weights AND normaliser state (incl. the running sums of torch_helpers.Normalizer)."""
    import functools

    import ceeus_b200

    # the reference calls torch.load(path) as torch 1.10 did; torch >= 2.6 defaults to weights_only=True, which refuses the numpy
    # arrays of its normaliser state dicts
    monkeypatch.setattr(torch, "load", functools.partial(torch.load, weights_only=False))
    env_ref = rh.make_construction_env(3)
    ref = rh.make_gnn_model(env_ref, hidden_dim=128, num_layers=2, running_stats=running, seed=3)
    rs = np.random.RandomState(0)
    for nz, dim in ((ref.input_normalizer_agent, 10), (ref.input_normalizer_obj_dyn, 12), (ref.action_normalizer, 4),
                    (ref.output_agent_normalizer, 10), (ref.output_obj_normalizer, 12)):
        nz.update(rs.standard_normal((50, dim)).astype(np.float32) * 0.3 + 0.1)
    env = ceeus_b200.construction_env(3)
    with tempfile.TemporaryDirectory() as d:
        p1, p2 = os.path.join(d, "ref.pt"), os.path.join(d, "b200.pt")
        ref.save(p1)
        m = ceeus_b200.B200GNNEnsemble(env=env, model_params=dict(n=5, hidden_dim=128, num_layers=2),
                                       normalize_w_running_stats=running)
        m.load(p1)
        for k, v in ref.model.state_dict().items():
            assert torch.equal(m.state_dict()[k], v.cpu())
        mean = ref.input_normalizer_obj_dyn.mean_tensor if running else ref.input_normalizer_obj_dyn.mean
        np.testing.assert_allclose(m._normalizers["in_dyn"][0], np.asarray(mean).reshape(-1), rtol=1e-6)
        m.save(p2)
        ref2 = rh.make_gnn_model(env_ref, hidden_dim=128, num_layers=2, running_stats=running, seed=9)
        ref2.load(p2, for_training=False)
        for k, v in ref.model.state_dict().items():
            assert torch.equal(ref2.model.state_dict()[k], v)
        a, b = ref.input_normalizer_obj_dyn.state_dict(), ref2.input_normalizer_obj_dyn.state_dict()
        assert set(a) == set(b)
        for k in a:
            np.testing.assert_allclose(np.asarray(b[k], np.float64), np.asarray(a[k], np.float64), rtol=1e-6, err_msg=k)
        saved = torch.load(p2, weights_only=False)
        assert set(saved) >= {"model", "input_normalizer_agent", "output_obj_normalizer"}


def test_sync_from_reference_model_object():
    import ceeus_b200

    env_ref = rh.make_playground_env(4)
    ref = rh.make_gnn_model(env_ref, hidden_dim=64, num_layers=2, running_stats=True, seed=5)
    m = ceeus_b200.B200GNNEnsemble(env=ceeus_b200.playground_env(4), model_params=dict(n=5, hidden_dim=64, num_layers=2),
                                   normalize_w_running_stats=True)
    v0 = m.weights_version
    m.sync_from(ref)
    assert m.weights_version > v0
    assert set(m.state_dict()) == set(ref.model.state_dict())
    assert m._normalizers["in_stat"][0].shape == (3,)
