#!/bin/bash
# usage: scripts/scale.sh N  -> gpurun_out/r2_scale_N.jsonl : BASELINE.json's multi-GPU configs at N GPUs of one node
N=$1; out=gpurun_out/r2_scale_$N.jsonl; : > $out
run() {  # args: bench.py flags
  if [ "$N" = 1 ]; then python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-per-config "$@" 2>/dev/null | tail -1 >> $out
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline "$@" 2>/dev/null | grep '^{' | tail -1 >> $out; fi
}
run --config c2                      # weak: 4096 trajectories per GPU, candidate all-gather inside the step's graph
run --config c3 --scaling strong     # 16384 trajectories in total (BASELINE configs[2])
run --config c4 --scaling strong     # 64 environments x 1024 in total, environments sharded, no collective (configs[3])
run --config c5                      # dense MLP, 8192 per GPU (configs[4])
python - <<PY
import json
for l in open("$out"):
    d = json.loads(l)
    print("$N GPUs", d["config"]["workload"][:34], d["scaling"], "global_traj", d["config"]["global_traj"], "ms/step %.3f" % d["ms_per_step"], "e2e %.3f" % d["e2e"]["ms_per_step"], "value %.3e" % d["value"], d.get("ranks_agree"))
PY
