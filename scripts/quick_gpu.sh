# quick GPU check: tensor-core parity tests + the headline bench lines (no CPU baseline)
timeout 600 python -m pytest tests/test_gpu_tc_parity.py -x -q 2>&1 | tail -5
for c in ${CONFIGS:-c2 c3 c4}; do
  timeout 300 python bench.py --config $c --steps 10 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$c', 'ms/step', round(d['ms_per_step'],3), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],4))"
done
