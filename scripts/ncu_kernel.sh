# one `ncu --set full` capture of a named kernel of a bench config: scripts/ncu_kernel.sh <config> <kernel-regex> <out-name>
CFG=${1:-c2}; KRN=${2:-cost_kernel}; OUT=${3:-prof}
ncu --set full --clock-control none --import-source on -k regex:$KRN -s 3 -c 1 -f -o gpurun_out/$OUT python bench.py --config $CFG --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${OUT}_ncu.log 2>&1
