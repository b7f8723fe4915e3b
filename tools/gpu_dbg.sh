#!/bin/bash
# usage: tools/gpu_dbg.sh <workload> "<debug values>" [extra env assignments]
W=$1; shift; DS=$1; shift
for d in $DS; do
  env "$@" LSM_DEBUG=$d python bench.py --workload $W --steps 60 --warmup 5 --no-cpu-baseline --e2e-steps 3 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; l=d['config']['launch']
print('$* LSM_DEBUG=$d', 'step_ms', round(d['ms_per_step'],4), 'b2b', round(d['ms_per_step_back_to_back'],4), 'emit_ms', round(r['mean_launch_ms'],4), l['emit_block_threads'], l['emit_regs_per_thread'], l['emit_blocks_per_sm'])"
done
