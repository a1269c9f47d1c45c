"""bench.py -- headline benchmark of the iCEM planning hot path (BASELINE.json metric: model-steps/s and ms per MPC step).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c2|c3|c5|c1|c4] [--precision fp32|tc]
    python bench.py --impl reference ...     # the reference's own CPU torch path on the host cores

A "step" is one MPC planning step (TorchMpcICem.get_action): opt_iterations x {sample, rollout, costs, top-k, refit}.
model-steps = trajectories x horizon x ensemble x iterations (kept elites are not re-simulated).
At N=1 the workload is BASELINE.json configs[1] (construction, 2 blocks, GNN ensemble of 5, 4096 trajectories, horizon 30,
3 iCEM iterations); at N>1 every GPU keeps 4096 trajectories (weak scaling) and each iteration exchanges the ranks' elite
candidates.  Synthetic data: seeded random-init weights, synthetic start observation (cee-us_b200/synthetic.py).
The same JSON line carries a `per_config` block with the other BASELINE configs (c1, c3, c4, c5) measured in the same run.

The GPU arm imports nothing from oracle/ or tests/.  Only the `cpu_baseline` legs and `--impl reference` execute the CPU
checker code: the reference's own classes from oracle/_ref/reference_mbrl.zip (kind "reference"; packed by build() from the
mounted reference tree) or, where that archive is missing, the restatement oracle/icem_oracle.py (kind "port").
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

E, ITERS = 5, 3
CONFIGS = {
    # name: (description, case dict).  P is per GPU (and per environment).
    "c1": ("construction N=2 GNN ensemble E=5 Hd=128, P=128, H=20, 3 iCEM iters, curiosity sum (BASELINE configs[0])",
           dict(env="construction", n_obj=2, model="gnn", hidden=128, layers=2, P=128, H=20, curiosity=True, cost="sum",
                sampler={})),
    "c2": ("construction N=2 GNN ensemble E=5 Hd=128, P=4096/GPU, H=30, 3 iCEM iters, curiosity sum (BASELINE configs[1])",
           dict(env="construction", n_obj=2, model="gnn", hidden=128, layers=2, P=4096, H=30, curiosity=True, cost="sum",
                sampler={})),
    "c3": ("construction N=6 (30 edges) GNN ensemble, stack task cost + disagreement, best, P=2048/GPU, H=30 (BASELINE configs[2])",
           dict(env="construction", n_obj=6, model="gnn", hidden=128, layers=2, P=2048, H=30, curiosity=True,
                extrinsic=True, cost="best", sampler=dict(init_std=0.5, use_mean_actions=False), strong_total_traj=16384)),
    "c4": ("playground N=4 GNN ensemble Hd=64, 8 envs x 1024 traj per GPU, H=30 (BASELINE configs[3])",
           dict(env="playground", n_obj=4, model="gnn", hidden=64, layers=2, P=1024, H=30, curiosity=True, cost="sum",
                num_envs=8, strong_total_envs=64,
                sampler=dict(keep_previous_elites=False, shift_elites_over_time=False, fraction_elites_reused=0.1))),
    "c5": ("construction N=4 dense-MLP ensemble 62-256x3-58 SiLU, P=8192/GPU, H=30 (BASELINE configs[4])",
           dict(env="construction", n_obj=4, model="mlp", hidden=256, layers=3, P=8192, H=30, curiosity=True, cost="sum",
                sampler={})),
}
SAMPLER = dict(opt_iterations=ITERS, elites_size=10, alpha=0.1, init_std=0.8, relative_init=True, execute_best_elite=True,
               keep_previous_elites=True, shift_elites_over_time=True, finetune_first_action=False,
               fraction_elites_reused=0.3, use_mean_actions=True, colored_noise=True, noise_beta=3.5,
               use_ensemble_cost_std=False)


def sampler_of(c):
    sp = dict(SAMPLER)
    sp.update(c.get("sampler", {}))
    return sp


# ---- workload (product-side synthetic inputs; shared by both arms) -------------------------------------------------
def build_workload(c):
    import ceeus_b200
    from ceeus_b200 import synthetic as syn

    env = ceeus_b200.construction_env(c["n_obj"]) if c["env"] == "construction" else ceeus_b200.playground_env(c["n_obj"])
    shapes = (syn.gnn_weight_shapes if c["model"] == "gnn" else syn.mlp_weight_shapes)(env, c["hidden"], c["layers"], E)
    weights = syn.synth_weights(shapes, 0)
    norms = syn.synth_normalizers(syn.normalizer_dims(env, c["model"]), 0, identity=True)
    obs, goal = syn.synth_start_obs(env, 1)
    flops = syn.flops_per_model_step(env, c["model"], c["hidden"], c["layers"])
    return dict(env=env, weights=weights, norms=norms, obs=obs, goal=goal, flops=flops)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(bf16_burst=d["bf16_tflops"], bf16_sustained=d["bf16_tflops_sustained"], hbm=d["hbm_gbs"], source="measured")
    return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, f"/tmp/ceeus_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in open(self.path):
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---- CPU arm: the reference's own classes (or the port) timed on the host cores ------------------------------------
def _reference_stepper(c, w, sample_P):
    """-> (kind, step_fn): one call of step_fn() = one MPC step (get_action) of the reference's CPU torch path."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_harness as rh

    sp = sampler_of(c)
    if rh.reference_available():
        if c["env"] == "construction":
            env = rh.make_construction_env(c["n_obj"], goal=w["goal"])
        else:
            env = rh.make_playground_env(c["n_obj"], goal=w["goal"])
        make = rh.make_gnn_model if c["model"] == "gnn" else rh.make_mlp_model
        model = make(env, hidden_dim=c["hidden"], num_layers=c["layers"], running_stats=False)
        rh.load_into_reference(model, w["weights"], w["norms"], False)
        ctrl = rh.make_controller(env, model, curiosity=c.get("curiosity", True), horizon=c["H"], num_traj=sample_P,
                                  cost_along_trajectory=c.get("cost", "sum"), sampler=sp,
                                  extrinsic_reward=c.get("extrinsic", False))
        ctrl.beginning_of_rollout(observation=w["obs"], state=None, mode="train")
        return "reference", lambda: ctrl.get_action(w["obs"], None)
    # the restatement (same torch ops in the same order as the reference), only where the archive is missing
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import icem_oracle as orc
    from cases import SeededNoise

    spec = orc.construction_spec(c["n_obj"]) if c["env"] == "construction" else orc.playground_spec(c["n_obj"])
    cls = orc.GNNEnsembleOracle if c["model"] == "gnn" else orc.MLPEnsembleOracle
    om = cls(spec=spec, weights=w["weights"], normalizers=w["norms"], hidden=c["hidden"], num_layers=c["layers"])
    kw = dict(horizon=c["H"], num_traj=sample_P, cost_along_trajectory=c.get("cost", "sum"),
              curiosity=c.get("curiosity", True), extrinsic_reward=c.get("extrinsic", False), stable_sort=False)
    kw.update(c.get("sampler", {}))
    oc = orc.ICEMOracle(om, orc.ICEMParams(**kw), SeededNoise(7, 3.5))
    oc.set_goal(w["goal"])
    oc.beginning_of_rollout()

    def step():
        oc.get_action(w["obs"])
        oc.trace.clear()

    return "port", step


def cpu_measure(c, w, steps, warm, budget_s):
    """Times `steps` MPC steps (after `warm` untimed ones) of the CPU path on all host threads, on the largest population
    P, P/2, P/4 ... >= 128 whose run is expected to fit `budget_s` (calibrated with one step at 128 trajectories); for the
    multi-environment config one environment is planned (the reference plans one environment at a time)."""
    import contextlib

    with contextlib.redirect_stdout(sys.stderr):  # the reference prints its evaluation count: keep stdout to the one JSON line
        return _cpu_measure(c, w, steps, warm, budget_s)


def _cpu_measure(c, w, steps, warm, budget_s):
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    cal_P = min(c["P"], 128)
    kind, step = _reference_stepper(c, w, cal_P)
    step()
    t0 = time.perf_counter()
    step()
    thr = cal_P * c["H"] * E * ITERS / (time.perf_counter() - t0)  # model-steps/s at the calibration size
    sample_P = c["P"]
    while sample_P > 128 and (steps + warm) * sample_P * c["H"] * E * ITERS / thr > budget_s:
        sample_P //= 2
    if sample_P != cal_P:
        kind, step = _reference_stepper(c, w, sample_P)
    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    ms = 1e3 * (time.perf_counter() - t0) / steps
    msteps = sample_P * c["H"] * E * ITERS
    nenv = c.get("num_envs", 1)
    sample = (f"{sample_P} of {c['P']} trajectories" + (f", 1 of {nenv} environments" if nenv > 1 else "") +
              f" per MPC step; {steps} timed steps after {warm} warm-up; "
              f"{'the unmodified reference classes (oracle/_ref/reference_mbrl.zip), its own CPU colored-noise sampler' if kind == 'reference' else 'oracle/icem_oracle.py restatement'}")
    return dict(value=msteps / (ms / 1e3), unit="model-steps/s", cores=torch.get_num_threads(), kind=kind, sample=sample,
                ms_per_step_sample=ms, sample_traj=sample_P, full_config=bool(sample_P == c["P"] and nenv == 1))


# ---- GPU arm ------------------------------------------------------------------------------------------------------
def shard(c, world, scaling):
    """-> (case dict for one rank, trajectories are sharded?).  weak: every GPU keeps the config's per-GPU population.
    strong: BASELINE.json's TOTAL population (c3: 16384 trajectories; c4: 64 environments x 1024; others: N=1 size) is split over
    the GPUs -- trajectories of ONE planning problem are sharded with a candidate all-gather per iteration, whole environments
    are sharded without any collective (each rank plans its own)."""
    c = dict(c)
    if scaling == "strong":
        if c.get("num_envs", 1) > 1:
            total_envs = c.get("strong_total_envs", c["num_envs"])
            assert total_envs % world == 0
            c["num_envs"] = total_envs // world
            return c, False
        total = c.get("strong_total_traj", c["P"])
        assert total % world == 0
        c["P"] = total // world
        return c, world > 1
    return c, world > 1 and c.get("num_envs", 1) == 1


def make_controller(c, w, world, precision, seed=1234):
    import ceeus_b200

    nenv = c.get("num_envs", 1)
    env = w["env"]
    env.goal = w["goal"].copy() if nenv == 1 else np.tile(w["goal"], (nenv, 1))
    if c["model"] == "gnn":
        model = ceeus_b200.B200GNNEnsemble(env=env, model_params=dict(n=E, hidden_dim=c["hidden"], num_layers=c["layers"],
                                                                       layer_norm=True, act_fn="relu"),
                                           normalize_w_running_stats=False, backend_params=dict(precision=precision))
    else:
        model = ceeus_b200.B200MLPEnsemble(env=env, model_params=dict(n=E, hidden_dim=c["hidden"], num_layers=c["layers"],
                                                                       act_fn="silu", layer_norm=False),
                                           backend_params=dict(precision=precision))
    model.load_state_dict({k: torch.from_numpy(v) for k, v in w["weights"].items()})
    model.set_normalizers(w["norms"])
    kw = dict(env=env, forward_model=model, horizon=c["H"], num_simulated_trajectories=c["P"] * world,
              action_sampler_params=sampler_of(c), cost_along_trajectory=c.get("cost", "sum"), logging=False,
              backend_params=dict(precision=precision, seed=seed, num_envs=nenv, distributed=world > 1))
    if c.get("curiosity", True):
        return ceeus_b200.B200CuriosityMpcICem(extrinsic_reward=c.get("extrinsic", False), **kw)
    return ceeus_b200.B200MpcICem(**kw)


def gpu_measure(name, steps, warmup, precision, world, rank, dev, flush, dist, scaling="weak"):
    """-> dict with the device-timed value, the end-to-end value through get_action (host obs in, host action out) and the
    roofline of the rollout kernel timed inside the steps.  All times are CUDA-event times, max over ranks."""
    desc, c = CONFIGS[name]
    c, traj_sharded = shard(c, world, scaling)
    w = build_workload(c)
    nenv = c.get("num_envs", 1)
    if not traj_sharded and world > 1:
        w["obs"] = w["obs"].copy()
        w["obs"][:2] += np.float32(0.01 * rank)  # env-sharded ranks plan different problems
    ctrl = make_controller(c, w, world if traj_sharded else 1, precision, seed=1234 + (0 if traj_sharded else rank))
    eng = ctrl.engine
    obs_b = w["obs"] if nenv == 1 else np.tile(w["obs"], (nenv, 1))
    ctrl.beginning_of_rollout(observation=obs_b, state=None, mode="train")
    eng.obs_dev_.copy_(torch.from_numpy(np.ascontiguousarray(obs_b.reshape(nenv, -1))))
    eng.set_goal(w["env"].goal)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, K):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        barrier()
        for a, b in ev:
            flush.fill_(1.0)  # evict L2 (126 MB) between timed steps; outside the timed interval
            a.record()
            fn()
            b.record()
        barrier()
        return sum(a.elapsed_time(b) for a, b in ev)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(warmup):
        ctrl.plan_on_device()
    l0 = eng.kernel_launches()
    total_ms = max_over_ranks(timed(ctrl.plan_on_device, steps))  # one CUDA-graph launch per step
    launches = eng.kernel_launches() - l0
    # the dominant kernel INSIDE the step: the same steps enqueued kernel by kernel with a CUDA-event pair around every rollout
    eng.profile_rollouts(True)
    prof_ms = max_over_ranks(timed(ctrl.plan_on_device, max(2, min(steps, 10))))
    roll = eng.profile_rollouts(False)
    # end to end through the public API: host obs -> get_action -> host action, every step
    for _ in range(2):
        ctrl.get_action(obs_b, None)
    e2e_ms = max_over_ranks(timed(lambda: ctrl.get_action(obs_b, None), steps))
    agree = None
    if world > 1 and traj_sharded:  # every rank must have planned the same action and the same refitted distribution
        d = 0.0
        for t in (eng.action_out_, eng.mean_, eng.std_):
            ref = t.clone()
            dist.broadcast(ref, src=0)
            d = max(d, float((t - ref).abs().max()))
        d = max_over_ranks(d)
        agree = {"max_abs_diff_action_mean_std": d, "ok": d == 0.0}

    msteps_gpu = nenv * c["P"] * c["H"] * E * ITERS
    ms_per_step = total_ms / steps
    kernel_ms = max_over_ranks(sum(roll) / max(1, len(roll)))
    peaks = measured_peaks()
    tc_mode = precision != "fp32"
    if tc_mode:  # fp16 operands on tcgen05: the dense bf16/fp16 peak; the kernel runs inside a longer step -> burst figure kept
        peak, peak_src = peaks["bf16_burst"], f"{peaks['source']}: bf16_tflops (burst; fp16 operands = same tensor rate)"
    else:        # fp32 FMA path: reported against the tf32-class tensor peak it is meant to be replaced by
        peak, peak_src = peaks["bf16_burst"] / 2.0, f"{peaks['source']}: bf16_tflops / 2 (tf32-class dense peak)"
    flop_per_launch = float(w["flops"]) * nenv * c["P"] * c["H"] * E
    achieved = flop_per_launch / (kernel_ms / 1e3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):  # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture
        traffic = json.load(open(tpath)).get(f"{name}:{precision}")
    out = {
        "value": msteps_gpu * world / (ms_per_step / 1e3), "ms_per_step": ms_per_step,
        "dtype": "f32" if not tc_mode else "f16 operands, f32 accumulate",
        "config": {"workload": desc, "global_traj": c["P"] * nenv * world, "traj_per_gpu": c["P"] * nenv, "horizon": c["H"],
                   "ensemble": E, "icem_iters": ITERS, "num_envs": nenv, "precision": precision,
                   "tc_row_layout": eng.tc_row_layout() if tc_mode else None,
                   "parallelism": (f"traj-sharded x{world} (per iteration the ranks' k best candidates are exchanged by NVLink peer-memory stores + flags, "
                                    "or one NCCL all-gather, inside the step's CUDA graph)"
                                   if traj_sharded else f"env-sharded x{world} (no collective)") if world > 1 else "single GPU",
                   "scaling": scaling,
                   "l2": "256 MiB buffer written between timed steps (outside the timed interval)"},
        "e2e": {"value": msteps_gpu * world / (e2e_ms / steps / 1e3), "unit": "model-steps/s", "ms_per_step": e2e_ms / steps,
                "h2d_bytes_per_step": int(obs_b.size * 4), "d2h_bytes_per_step": int(nenv * eng.U * 4)},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "kernel": f"rollout_{c['model']}_{'tc' if tc_mode else 'simt'}_kernel",
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                     "peak_source": peak_src, "kernel_ms": kernel_ms, "kernel_launches_timed": len(roll),
                     "kernel_timing": "CUDA-event pair around every rollout launch inside MPC steps enqueued kernel by kernel "
                                      "right after the graph-launched timed steps (same stream, same buffers)",
                     "ms_per_step_profiled": prof_ms / max(2, min(steps, 10)),
                     "kernel_share_of_step": ITERS * kernel_ms / ms_per_step, "flop_per_launch": flop_per_launch},
    }
    if agree is not None:
        out["ranks_agree"] = agree
    ctrl.engine.close()
    return out, c, w


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--precision", default=os.environ.get("CEEUS_PRECISION", "auto"),
                    help="tc (tcgen05, fp16 operands / fp32 accumulate, parity 1e-3) | fp32 (parity 1e-5) | auto = tc")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-per-config", action="store_true")
    ap.add_argument("--cpu-budget-s", type=float, default=0.0, help="wall-clock budget of a CPU leg (0 = default)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="strong: BASELINE.json's total population (c3: 16384 trajectories, c4: 64 environments) split over the GPUs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.precision == "auto":
        args.precision = "tc"

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    desc, c = CONFIGS[args.config]
    nenv = c.get("num_envs", 1)
    base = {"metric": "model_steps_per_sec", "unit": "model-steps/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "data": "synthetic (seeded random-init weights, synthetic start observation, on-device Philox noise)"}

    # ---------------------------------------------------------------------------------------------------
    if args.impl == "reference":
        if rank != 0:
            return
        w = build_workload(c)
        warm = min(args.warmup, 1)
        r = cpu_measure(c, w, args.steps, warm, args.cpu_budget_s or 150.0)
        out = dict(base)
        out.update({"impl": "reference", "value": r["value"], "ms_per_step": r["ms_per_step_sample"], "warmup": warm,
                    "dtype": "f32", "data": "synthetic (same seeded weights / start observation as the GPU arm)",
                    "config": {"workload": desc, "traj_per_step_timed": r["sample_traj"], "traj_per_gpu": c["P"] * nenv,
                               "horizon": c["H"], "ensemble": E, "icem_iters": ITERS,
                               "note": "ms_per_step is the measured time of one MPC step over traj_per_step_timed trajectories "
                                       "(no extrapolation); at N > 1 the arm still times one GPU's share on the host cores"},
                    "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                    "e2e": {"value": r["value"], "unit": "model-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                    "gpu_launches": 0})
        print(json.dumps(out))
        return

    # ---------------------------------------------------------------------------------------------------
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev, dtype=torch.float32)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    main_out, c, w = gpu_measure(args.config, args.steps, args.warmup, args.precision, world, rank, dev, flush, dist,
                                 args.scaling)
    per = {}
    if world == 1 and not args.no_per_config:
        for name in sorted(CONFIGS):
            if name == args.config:
                continue
            o, cc, ww = gpu_measure(name, max(3, min(args.steps, 10)), 3, args.precision, 1, 0, dev, flush, dist)
            per[name] = {"workload": CONFIGS[name][0], "value": o["value"], "ms_per_step": o["ms_per_step"],
                         "e2e": o["e2e"]["value"], "e2e_ms_per_step": o["e2e"]["ms_per_step"], "frac": o["roofline"]["frac"],
                         "kernel_ms": o["roofline"]["kernel_ms"], "kernel_share_of_step": o["roofline"]["kernel_share_of_step"],
                         "achieved_tflops": o["roofline"]["achieved"], "gpu_launches": o["gpu_launches"]}
            per[name]["_inputs"] = (cc, ww)
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    out = dict(base)
    out.update(main_out)
    out["scaling"] = args.scaling
    out["clocks"] = clocks
    if not args.no_cpu_baseline:
        r = cpu_measure(c, w, 3, 1, args.cpu_budget_s or 25.0)
        out["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "ms_per_step_sample")}
    for name, p in per.items():
        cc, ww = p.pop("_inputs")
        if not args.no_cpu_baseline:
            r = cpu_measure(cc, ww, 2, 1, args.cpu_budget_s or 10.0)
            p["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
            p["e2e_over_cpu"] = p["e2e"] / r["value"]
    if per:
        out["per_config"] = per
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
