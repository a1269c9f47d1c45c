"""CPU: host logic, the C ABI's export table and struct layout, error behaviour without a GPU, and the world_size=2
(gloo) exchange of elite candidates."""
import ctypes
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ceeus_b200.h")


@pytest.fixture(scope="session")
def built_lib():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g

    g.build()
    import ceeus_b200

    return ceeus_b200._lib.load()


def test_library_exports_every_declared_symbol(built_lib):
    import ceeus_b200

    src = open(HEADER).read()
    declared = set(re.findall(r"\b(ceeus_[a-z0-9_]+)\s*\(", src))
    assert declared, "no declarations parsed"
    assert declared == set(ceeus_b200._lib.SIGNATURES), declared ^ set(ceeus_b200._lib.SIGNATURES)
    for name in declared:
        assert hasattr(built_lib, name), name
    assert built_lib.ceeus_abi_version() == ceeus_b200._lib.ABI_VERSION == 3


def test_ctypes_structs_match_c_layout(built_lib):
    import ceeus_b200

    fields_cfg = [f[0] for f in ceeus_b200._lib.CeeusConfig._fields_]
    fields_buf = [f[0] for f in ceeus_b200._lib.CeeusBuffers._fields_]
    prog = '#include <stdio.h>\n#include <stddef.h>\n#include "ceeus_b200.h"\nint main(){\n'
    prog += 'printf("%zu %zu\\n", sizeof(CeeusConfig), sizeof(CeeusBuffers));\n'
    for f in fields_cfg:
        prog += f'printf("%zu\\n", offsetof(CeeusConfig, {f}));\n'
    for f in fields_buf:
        prog += f'printf("%zu\\n", offsetof(CeeusBuffers, {f}));\n'
    prog += "return 0;}\n"
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(prog)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")])
        out = subprocess.check_output([os.path.join(d, "t")]).decode().split()
    assert int(out[0]) == ctypes.sizeof(ceeus_b200._lib.CeeusConfig)
    assert int(out[1]) == ctypes.sizeof(ceeus_b200._lib.CeeusBuffers)
    offs = [int(x) for x in out[2:]]
    want = [getattr(ceeus_b200._lib.CeeusConfig, f).offset for f in fields_cfg] + \
           [getattr(ceeus_b200._lib.CeeusBuffers, f).offset for f in fields_buf]
    assert offs == want


def test_no_cpu_fallback(built_lib):
    """Without a CUDA device the product refuses to run instead of silently computing on the host."""
    import ceeus_b200

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    env = ceeus_b200.construction_env(2)
    model = ceeus_b200.B200GNNEnsemble(env=env, model_params=dict(n=5, hidden_dim=128, num_layers=2))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ceeus_b200.Engine(horizon=5, num_traj=8, **model.engine_kwargs())
    cfg = ceeus_b200._lib.CeeusConfig()
    h = ctypes.c_void_p()
    assert built_lib.ceeus_create(ctypes.byref(cfg), ctypes.byref(h)) == ceeus_b200._lib.ERR_INVALID
    assert b"abi_version" in built_lib.ceeus_last_error(None)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "cee-us_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "icem_oracle" not in txt and "import oracle" not in txt and "ref_harness" not in txt, f


def test_env_description_and_registry(built_lib):
    import ceeus_b200
    from ceeus_b200 import registry
    from ceeus_b200.engine import weight_key_order

    d = ceeus_b200.describe_env(ceeus_b200.construction_env(6))
    assert (d["obs_dim"], d["state_dim"], d["goal_dim"], d["action_dim"]) == (103, 82, 21, 4)
    assert d["threshold"] == 0.02 and d["task_cost_supported"]
    d = ceeus_b200.describe_env(ceeus_b200.playground_env(4))
    assert (d["obs_dim"], d["state_dim"], d["goal_dim"], d["action_dim"]) == (48, 40, 8, 2)
    f = ceeus_b200.describe_env(ceeus_b200.construction_env(2, case="Flip"))  # fpp_construction_env.py:93-99
    assert f["task_cost_supported"] and f["cost_case"] == ceeus_b200._lib.CASE_FLIP
    assert f["coord_weights"].tolist() == [1.0, 0.0, 0.0] and np.allclose(f["cost_thres"], [0.087, 0, 0])
    assert np.allclose(f["buffer_xyz"], 1e-4) and np.allclose(f["gripper_init"], [1.344, 0.749, 0.5319])
    sl = ceeus_b200.describe_env(ceeus_b200.construction_env(3, case="Slide", shaped_reward=False))  # :100-105
    assert sl["cost_case"] == ceeus_b200._lib.CASE_SLIDE and np.allclose(sl["buffer_xyz"], [0.1, 0.1, 0.42])
    assert np.allclose(sl["cost_thres"], [0.1, 0.1, 0.42]) and sl["threshold"] == 0.1
    assert registry.controller_from_string("mpc-curiosity-icem-b200") is ceeus_b200.B200CuriosityMpcICem
    assert registry.forward_model_from_string("ParallelGNNDeterministicEnsembleB200") is ceeus_b200.B200GNNEnsemble
    with pytest.raises(ImportError):
        registry.controller_from_string("nope")
    keys = weight_key_order("gnn", 2, True)
    assert len(keys) == 30 and keys[0] == "edge_mlp.layers.0.0.weight" and keys[-1] == "global_mlp.layers.2.0.bias"
    assert weight_key_order("mlp", 3, False) == [f"layers.{l}.0.{k}" for l in range(4) for k in ("weight", "bias")]


def test_model_dropin_hand_off_surface(built_lib, tmp_path):
    import ceeus_b200
    import icem_oracle as orc

    env = ceeus_b200.construction_env(2)
    m = ceeus_b200.B200GNNEnsemble(env=env, model_params=dict(n=5, hidden_dim=128, num_layers=2, layer_norm=True,
                                                               num_message_passing=1, act_fn="relu", output_act_fn="none"),
                                   train_params=dict(batch_size=125, epochs=25, iterations=0, learning_rate=1e-5),
                                   use_input_normalization=True, use_output_normalization=True, target_is_delta=True,
                                   normalize_w_running_stats=False)
    spec = orc.construction_spec(2)
    w = orc.synth_weights(orc.gnn_weight_shapes(spec, 128, 2, 5), 0)
    m.load_state_dict(w)
    v0 = m.weights_version
    m.set_normalizers(orc.synth_normalizers(orc.gnn_normalizer_dims(spec), 0, identity=False))
    assert m.weights_version == v0 + 1 and m.ensemble_size == 5
    path = str(tmp_path / "ckpt")
    m.save(path)
    m2 = ceeus_b200.B200GNNEnsemble(env=env, model_params=dict(n=5, hidden_dim=128, num_layers=2))
    m2.load(path)
    for k in w:
        assert torch.equal(m2.state_dict()[k], m.state_dict()[k])
    np.testing.assert_array_equal(m2._normalizers["out_dyn"][1], m._normalizers["out_dyn"][1])
    assert m.reset(np.zeros(43, np.float32))["obs"].shape == (43,)
    assert m.got_actual_observation_and_env_state(observation=None) is None
    # like the reference, the global-node variant accepts num_message_passing but always does one round (gnn_ensemble.py:100-114)
    mv = ceeus_b200.B200GNNEnsemble(env=env, model_params=dict(n=5, hidden_dim=128, num_message_passing=2))
    assert mv.engine_kwargs()["gnn_variant"] == 0 and "num_message_passing" not in mv.engine_kwargs()
    mh = ceeus_b200.B200GNNEnsemble(env=env, agent_as_global_node=False,
                                    model_params=dict(n=5, hidden_dim=128, num_message_passing=2, embedding_dim=32))
    assert mh.engine_kwargs()["gnn_variant"] == 1 and mh.engine_kwargs()["num_message_passing"] == 2
    with pytest.raises(ValueError):
        ceeus_b200.B200GNNEnsemble(env=env, agent_as_global_node=False, model_params=dict(n=5, hidden_dim=128))


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    import ceeus_b200  # noqa: F401
    from ceeus_b200.distributed import gather_candidates, merge_candidates_reference, shard_range

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    P, k, E, HU = 64, 5, 3, 8
    rec = 2 + E + HU
    g = torch.Generator().manual_seed(0)
    costs = torch.randn(P, generator=g)
    costs[10] = costs[40]  # a tie across shards: lower global index must win
    payload = torch.randn(P, E + HU, generator=g)
    lo, hi = shard_range(P, world, rank)
    local = costs[lo:hi]
    order = sorted(range(hi - lo), key=lambda i: (float(local[i]), i))[:k]
    cand = torch.zeros(1, k, rec)
    for j, i in enumerate(order):
        cand[0, j, 0] = local[i]
        cand[0, j, 1] = torch.tensor([lo + i], dtype=torch.int32).view(torch.float32)[0]
        cand[0, j, 2:] = payload[lo + i]
    allc = gather_candidates(cand)
    assert allc.shape == (world, 1, k, rec)
    recs, idx = merge_candidates_reference(allc[:, 0], k)
    want = sorted(range(P), key=lambda i: (float(costs[i]), i))[:k]
    ok = idx.tolist() == want and torch.equal(recs[:, 2:], payload[want])
    q.put((rank, ok, idx.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_topk_exchange_world_size_2_gloo(built_lib):
    import socket

    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert res[0][2] == res[1][2]  # identical elites on both ranks


def test_shard_range():
    import ceeus_b200  # noqa: F401
    from ceeus_b200.distributed import shard_range

    assert [shard_range(16384, 8, r) for r in (0, 7)] == [(0, 2048), (14336, 16384)]
    with pytest.raises(ValueError):
        shard_range(10, 4, 0)
