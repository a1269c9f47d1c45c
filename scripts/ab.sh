#!/bin/bash
# usage: scripts/ab.sh "<label>" cfg [env assignments...]  -> one line: label cfg ms_per_step e2e_ms kernel_ms launches frac
label=$1; cfg=$2; shift 2
env "$@" python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline --no-per-config 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$label', '$cfg', 'step %.3f e2e %.3f kernel %.3f launches %d frac %.3f' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['gpu_launches'], d['roofline']['frac']))"
