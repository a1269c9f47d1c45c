set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for c in c1 c2 c3 c4 c5; do timeout 600 python bench.py --config $c --steps 10 --warmup 5 > gpurun_out/bench_r1_$c.json 2>gpurun_out/bench_r1_$c.err; tail -c 300 gpurun_out/bench_r1_$c.json; done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1_reference.json 2>&1; tail -c 400 gpurun_out/bench_r1_reference.json
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_ll.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_tc.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_ll.log 2>&1
