"""bench.py -- headline benchmark of the iCEM planning hot path (BASELINE.json metric: model-steps/s and ms per MPC step).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c2|c3|c5|c1|c4] [--precision fp32|tc]
    python bench.py --impl reference ...        # the reference's CPU torch path (oracle port) on the host cores

A "step" is one MPC planning step (TorchMpcICem.get_action): opt_iterations x {sample, rollout, costs, top-k, refit}.
model-steps = trajectories x horizon x ensemble x iterations (kept elites are not re-simulated).
At N=1 the workload is BASELINE.json configs[1] (construction, 2 blocks, GNN ensemble of 5, 4096 trajectories,
horizon 30, 3 iCEM iterations); at N>1 every GPU keeps 4096 trajectories (weak scaling) and each iteration adds one
all-gather of the elite candidates.  Synthetic data: seeded random-init weights, synthetic start observation.

Only the `cpu_baseline` leg and `--impl reference` execute anything under oracle/ (as the thing being timed on the
CPU); the GPU arm never touches it.
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
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

CONFIGS = {
    # name: (description, case dict).  P is per GPU.
    "c1": ("construction N=2 GNN ensemble E=5 Hd=128, P=128, H=20, 3 iCEM iters, curiosity sum (BASELINE configs[0])",
           dict(env="construction", n_obj=2, model="gnn", hidden=128, layers=2, P=128, H=20, curiosity=True, cost="sum",
                sampler={})),
    "c2": ("construction N=2 GNN ensemble E=5 Hd=128, P=4096/GPU, H=30, 3 iCEM iters, curiosity sum (BASELINE configs[1])",
           dict(env="construction", n_obj=2, model="gnn", hidden=128, layers=2, P=4096, H=30, curiosity=True, cost="sum",
                sampler={})),
    "c3": ("construction N=6 (30 edges) GNN ensemble, stack task cost + disagreement, best, P=2048/GPU, H=30 (BASELINE configs[2])",
           dict(env="construction", n_obj=6, model="gnn", hidden=128, layers=2, P=2048, H=30, curiosity=True,
                extrinsic=True, cost="best", sampler=dict(init_std=0.5, use_mean_actions=False))),
    "c4": ("playground N=4 GNN ensemble Hd=64, 8 envs x 1024 traj per GPU, H=30 (BASELINE configs[3])",
           dict(env="playground", n_obj=4, model="gnn", hidden=64, layers=2, P=1024, H=30, curiosity=True, cost="sum",
                num_envs=8, sampler=dict(keep_previous_elites=False, shift_elites_over_time=False,
                                         fraction_elites_reused=0.1))),
    "c5": ("construction N=4 dense-MLP ensemble 62-256x3-58 SiLU, P=8192/GPU, H=30 (BASELINE configs[4])",
           dict(env="construction", n_obj=4, model="mlp", hidden=256, layers=3, P=8192, H=30, curiosity=True, cost="sum",
                sampler={})),
}


def case_dict(name, world):
    desc, c = CONFIGS[name]
    c = dict(c)
    c.update(kind="mpc", wseed=0, identity_norm=True, running=False, oseed=1, nseed=2)
    c.setdefault("sampler", {})
    return desc, c


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(bf16_burst=d["bf16_tflops"], bf16_sustained=d["bf16_tflops_sustained"], hbm=d["hbm_gbs"], source="measured")
    return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback")


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


def cpu_oracle_run(c, spec_inputs, sample_P, steps, warm):
    """Times the oracle port (torch CPU, all host threads) on a bounded sample of the same workload."""
    import icem_oracle as orc
    from helpers import SeededNoise, oracle_model, oracle_params

    spec, weights, norms, obs, goal = spec_inputs
    cc = dict(c)
    cc["P"] = sample_P
    params = oracle_params(cc)
    params.stable_sort = False  # reference default: fully_deterministic=False
    oc = orc.ICEMOracle(oracle_model(cc, spec, weights, norms), params, SeededNoise(7, 3.5))
    oc.set_goal(goal)
    oc.beginning_of_rollout()
    times = []
    for s in range(warm + steps):
        t0 = time.perf_counter()
        oc.get_action(obs)
        dt = time.perf_counter() - t0
        oc.trace.clear()
        if s >= warm:
            times.append(dt)
    ms = 1e3 * statistics.median(times)
    msteps = sample_P * cc["H"] * 5 * params.opt_iterations
    return dict(ms_per_step=ms, value=msteps / (ms / 1e3), model_steps=msteps)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--precision", default=os.environ.get("CEEUS_PRECISION", "auto"),
                    help="tc (tcgen05, fp16 operands / fp32 accumulate, parity 1e-3) | fp32 (parity 1e-5) | auto = tc")
    ap.add_argument("--cpu-sample-traj", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from cases import build_inputs
    import icem_oracle as orc_meta  # metadata only (flops table); never on the GPU arm's compute path

    desc, c = case_dict(args.config, world)
    if args.precision == "auto":
        args.precision = "tc"
    inputs = build_inputs(c)
    spec, weights, norms, obs, goal = inputs
    E, iters = 5, 3
    nenv = c.get("num_envs", 1)
    model_steps_per_gpu = nenv * c["P"] * c["H"] * E * iters
    from helpers import oracle_model

    flops_ms = orc_meta.flops_per_model_step(oracle_model(c, spec, weights, norms))
    base = {"metric": "model_steps_per_sec", "unit": "model-steps/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "data": "synthetic (seeded random-init weights, synthetic start observation, on-device Philox noise)"}

    # ---------------------------------------------------------------------------------------------------
    if args.impl == "reference":
        if rank != 0:
            return
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        sample_P = args.cpu_sample_traj or (256 if c["n_obj"] <= 2 else 64 if c["model"] == "gnn" else 512)
        r = cpu_oracle_run(c, inputs, sample_P, max(1, min(args.steps, 3)), min(args.warmup, 1))
        cb = {"value": r["value"], "unit": "model-steps/s", "cores": torch.get_num_threads(), "kind": "port",
              "sample": f"{sample_P} of {c['P']} trajectories per MPC step (throughput is flat in P, SURVEY.md section 6), "
                        f"median of {max(1, min(args.steps, 3))} steps"}
        out = dict(base)
        out.update({"impl": "reference", "value": r["value"], "ms_per_step": r["ms_per_step"] * (c["P"] * nenv / sample_P) * args.gpus,
                    "dtype": "f32", "config": {"workload": desc, "global_traj": c["P"] * nenv * args.gpus,
                                               "horizon": c["H"], "ensemble": E, "icem_iters": iters,
                                               "note": "reference CPU torch path restated in oracle/icem_oracle.py (pinned "
                                                       "to the reference by tests/golden); Python reference cannot travel "
                                                       "to the GPU box"},
                    "cpu_baseline": cb,
                    "e2e": {"value": r["value"], "unit": "model-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                    "gpu_launches": 0})
        print(json.dumps(out))
        return

    # ---------------------------------------------------------------------------------------------------
    import torch.distributed as dist
    from helpers import product_controller, product_env, product_model

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cg = dict(c)
    cg["P"] = c["P"] * world  # global population; the controller shards it
    env = product_env(cg, goal if nenv == 1 else np.tile(goal, (nenv, 1)))
    model = product_model(cg, env, weights, norms, precision=args.precision)
    ctrl = product_controller(cg, env, model, precision=args.precision, seed=1234, num_envs=nenv)
    eng = ctrl.engine
    obs_b = obs if nenv == 1 else np.tile(obs, (nenv, 1))
    ctrl.beginning_of_rollout(observation=obs_b, state=None, mode="train")
    eng.obs_dev_.copy_(torch.from_numpy(np.ascontiguousarray(obs_b.reshape(nenv, -1))))
    eng.set_goal(env.goal)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev, dtype=torch.float32)

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

    for _ in range(args.warmup):
        ctrl.plan_on_device()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = eng.kernel_launches()
    total_ms = max_over_ranks(timed(ctrl.plan_on_device, args.steps))
    launches = eng.kernel_launches() - l0
    # end to end through the public API: host obs -> get_action -> host action, every step
    for _ in range(2):
        ctrl.get_action(obs_b, None)
    e2e_ms = max_over_ranks(timed(lambda: ctrl.get_action(obs_b, None), args.steps))
    # dominant kernel alone (same stream, same inputs as the steps above): the rollout launch
    start = eng.obs_dev_
    acts = eng.actions_
    nroll = 3 * max(2, min(args.steps, 10))
    roll_ms = max_over_ranks(timed(lambda: eng.rollout(start, acts), nroll)) / nroll
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_per_step = total_ms / args.steps
    value = model_steps_per_gpu * world / (ms_per_step / 1e3)
    e2e_value = model_steps_per_gpu * world / (e2e_ms / args.steps / 1e3)
    peaks = measured_peaks()
    tc_mode = args.precision != "fp32"
    if tc_mode:  # fp16 operands on tcgen05: the dense bf16/fp16 peak; the rollout kernel is timed alone -> burst figure
        tc_peak, peak_src = peaks["bf16_burst"], f"{peaks['source']}: bf16_tflops (burst; fp16 operands, kernel timed alone)"
    else:        # fp32 FMA path: reported against the tf32-class tensor peak it is meant to be replaced by
        tc_peak, peak_src = peaks["bf16_burst"] / 2.0, f"{peaks['source']}: bf16_tflops / 2 (tf32-class dense peak)"
    flop_per_launch = float(flops_ms) * nenv * c["P"] * c["H"] * E
    achieved = flop_per_launch / (roll_ms / 1e3) / 1e12
    kernel_name = f"rollout_{c['model']}_{'tc' if tc_mode else 'simt'}_kernel"
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):  # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture
        traffic = json.load(open(tpath)).get(f"{args.config}:{args.precision}")
    out = dict(base)
    out.update({
        "value": value, "ms_per_step": ms_per_step, "dtype": "f32" if not tc_mode else "f16 operands, f32 accumulate",
        "config": {"workload": desc, "global_traj": c["P"] * nenv * world, "traj_per_gpu": c["P"] * nenv, "horizon": c["H"],
                   "ensemble": E, "icem_iters": iters, "num_envs": nenv, "precision": args.precision,
                   "parallelism": f"traj-sharded x{world}", "l2": "256 MiB buffer written between timed steps (outside the timed interval)"},
        "e2e": {"value": e2e_value, "unit": "model-steps/s", "ms_per_step": e2e_ms / args.steps,
                "h2d_bytes_per_step": int(obs_b.size * 4), "d2h_bytes_per_step": int(nenv * spec.action_dim * 4)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": kernel_name,
                     "achieved": achieved, "peak": tc_peak, "unit": "TFLOP/s", "frac": achieved / tc_peak, "traffic": traffic,
                     "peak_source": peak_src,
                     "kernel_ms": roll_ms, "kernel_share_of_step": min(1.0, 3 * roll_ms / ms_per_step),
                     "flop_per_launch": flop_per_launch},
    })
    if not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        sample_P = args.cpu_sample_traj or (256 if c["n_obj"] <= 2 else 64 if c["model"] == "gnn" else 512)
        r = cpu_oracle_run(c, inputs, sample_P, 3, 1)
        out["cpu_baseline"] = {"value": r["value"], "unit": "model-steps/s", "cores": torch.get_num_threads(), "kind": "port",
                               "ms_per_step_sample": r["ms_per_step"],
                               "sample": f"{sample_P} of {c['P']} trajectories per MPC step, median of 3 steps after 1 warm-up"}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
