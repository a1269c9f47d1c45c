# Round measurement on one B200: GPU tests, the five BASELINE configs, the reference arm, the launch list of the default
# bench command and one `ncu --set full` capture of the rollout kernel (configs c2 and c3).  Outputs under gpurun_out/.
set -x
R=${ROUND:-r1}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for c in c1 c2 c3 c4 c5; do timeout 600 python bench.py --config $c --steps 10 --warmup 5 > gpurun_out/bench_${R}_$c.json 2>gpurun_out/bench_${R}_$c.err; tail -c 300 gpurun_out/bench_${R}_$c.json; done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${R}_reference.json 2>&1; tail -c 400 gpurun_out/bench_${R}_reference.json
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_ll.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${R}_tc.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_ll.log 2>&1
bash scripts/ncu_top.sh c2 prof_${R}_tc_c2
bash scripts/ncu_top.sh c3 prof_${R}_tc_c3
bash scripts/ncu_top.sh c5 prof_${R}_tc_c5
