# one `ncu --set full` capture of the rollout kernel of a bench config: scripts/ncu_top.sh <config> <out-name>
CFG=${1:-c2}; OUT=${2:-prof}
python bench.py --config $CFG --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${OUT}_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:rollout_ -s 3 -c 1 -f -o gpurun_out/$OUT python bench.py --config $CFG --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${OUT}_ncu.log 2>&1
