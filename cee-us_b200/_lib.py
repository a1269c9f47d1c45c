"""ctypes binding of libceeus_b200.so (include/ceeus_b200.h).  There is no fallback: if the CUDA library is
missing or fails to load, importing the product path raises."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CEEUS_LIB") or os.path.join(HERE, "libceeus_b200.so")  # CEEUS_LIB: development builds (A/B runs)

ABI_VERSION = 3
OK, ERR_INVALID, ERR_CUDA, ERR_STATE = 0, -1, -2, -3
MODEL_GNN, MODEL_MLP = 0, 1
ACT = {"none": 0, None: 0, "relu": 1, "silu": 2, "swish": 2, "tanh": 3}
# "tc" = tcgen05 tensor cores with fp16 operands (10-bit mantissa, the TF32 mantissa width) and fp32 accumulation; it is NOT
# TF32 (5-bit exponent: the kernels saturate and raise a range flag, ceeus_tc_status), so no "tf32" alias is offered.
PREC = {"fp32": 0, "tc": 1, "fp16": 1}
ENV_NONE, ENV_CONSTRUCTION, ENV_PLAYGROUND, ENV_EXTERNAL, ENV_ROBODESK = 0, 1, 2, 3, 4
ROBODESK_TASKS = ("open_slide", "open_drawer", "push_red", "push_green", "push_blue", "ball_off_table",
                  "upright_block_off_table", "flat_block_off_table")  # CEEUS_RD_*
CASE_TOWER, CASE_PICKANDPLACE, CASE_FLIP, CASE_SLIDE = 0, 1, 2, 3
COST = {"sum": 0, "best": 1}


class CeeusConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "abi_version", "model_kind", "ensemble_size", "n_obj", "agent_dim", "dyn_dim", "stat_dim", "action_dim",
        "goal_dim", "hidden_dim", "num_layers", "act_fn", "layer_norm", "aggr_mean", "precision",
        "num_envs", "num_traj", "horizon", "opt_iterations", "num_elites", "num_kept", "relative_init",
        "use_mean_actions", "keep_previous_elites", "shift_elites_over_time", "execute_best_elite",
        "use_ensemble_cost_std", "colored_noise", "cost_along_trajectory", "curiosity", "extrinsic_reward")] + [
        ("alpha", C.c_float), ("init_std", C.c_float), ("noise_beta", C.c_float), ("extrinsic_scale", C.c_float),
        ("env_kind", C.c_int32), ("cost_case", C.c_int32), ("sparse", C.c_int32), ("shaped_reward", C.c_int32),
        ("threshold", C.c_float), ("buffer_threshold", C.c_float),
        ("coord_weights", C.c_float * 3), ("buffer_xyz", C.c_float * 3), ("cost_thres", C.c_float * 3),
        ("gripper_init", C.c_float * 3),
        ("world_size", C.c_int32), ("rank", C.c_int32), ("cost_in_rollout_tail", C.c_int32),
        ("gnn_variant", C.c_int32), ("edge_global", C.c_int32), ("num_message_passing", C.c_int32),
        ("embedding_dim", C.c_int32), ("tc_row_layout", C.c_int32),
        ("seed", C.c_uint64)]


_FP = C.POINTER(C.c_float)


class CeeusBuffers(C.Structure):
    _fields_ = [("mean", C.c_void_p), ("std", C.c_void_p), ("actions", C.c_void_p), ("traj", C.c_void_p),
                ("costs_per_model", C.c_void_p), ("costs", C.c_void_p), ("elite_actions", C.c_void_p),
                ("elite_costs", C.c_void_p), ("elite_cpm", C.c_void_p), ("elite_idx", C.c_void_p),
                ("cand", C.c_void_p), ("action_out", C.c_void_p), ("obs", C.c_void_p)]


# name -> (restype, argtypes); must list every symbol include/ceeus_b200.h declares (tests check the export table)
SIGNATURES = {
    "ceeus_abi_version": (C.c_int, []),
    "ceeus_last_error": (C.c_char_p, [C.c_void_p]),
    "ceeus_create": (C.c_int, [C.POINTER(CeeusConfig), C.POINTER(C.c_void_p)]),
    "ceeus_destroy": (None, [C.c_void_p]),
    "ceeus_num_weight_tensors": (C.c_int, [C.c_void_p]),
    "ceeus_weight_tensor_shape": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "ceeus_set_weights": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_void_p]),
    "ceeus_set_normalizers": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "ceeus_set_action_bounds": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "ceeus_set_goal": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ceeus_bind_buffers": (C.c_int, [C.c_void_p, C.POINTER(CeeusBuffers)]),
    "ceeus_cand_record_len": (C.c_int, [C.c_void_p]),
    "ceeus_rollout": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "ceeus_costs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ceeus_sample": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_uint32,
                               C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
    "ceeus_begin_rollout": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ceeus_plan_iter_local": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ceeus_plan_iter_rollout": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ceeus_plan_iter_costs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "ceeus_plan_iter_update": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "ceeus_plan_finish": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ceeus_plan_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ceeus_plan_step_resident": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "ceeus_p2p_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ceeus_p2p_import": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ceeus_comm_unique_id": (C.c_int, [C.c_void_p]),
    "ceeus_comm_init": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ceeus_kernel_launches": (C.c_int64, [C.c_void_p]),
    "ceeus_has_elites": (C.c_int, [C.c_void_p]),
    "ceeus_profile_rollouts": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_float), C.c_int]),
    "ceeus_tc_status": (C.c_int, [C.c_void_p]),
    "ceeus_tc_tiling": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "ceeus_tc_row_layout": (C.c_int, [C.c_void_p]),
}

_lib = None


def load():
    """Loads the shared library (once).  Raises if it has not been built: no CPU fallback exists."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). ceeus_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        if lib.ceeus_abi_version() != ABI_VERSION:
            raise RuntimeError("libceeus_b200.so ABI version mismatch; rebuild")
        _lib = lib
    return _lib


# Diagnostics outside the product ABI (csrc/debug_tools.h): libceeus_b200_probe.so, test infrastructure
PROBE_PATH = os.path.join(HERE, "libceeus_b200_probe.so")
PROBE_SIGNATURES = {
    "ceeus_selftest_tmem": (C.c_int, [C.c_void_p]),
    "ceeus_selftest_tmem_contention": (C.c_int, [C.c_int, C.c_void_p]),
    "ceeus_selftest_umma": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                      C.POINTER(C.c_int)]),
}
_probe = None


def load_probe():
    global _probe
    if _probe is None:
        lib = C.CDLL(PROBE_PATH)
        for name, (res, args) in PROBE_SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _probe = lib
    return _probe


def check(rc, handle=None):
    if rc != 0:
        msg = load().ceeus_last_error(handle)
        raise RuntimeError(f"libceeus_b200 error {rc}: {msg.decode() if msg else '?'}")
