#!/bin/bash
for w in 1 2 4; do for d in 0 32; do
  LSM_WPE=$w LSM_DEBUG=$d python bench.py --workload cfg2 --steps 60 --warmup 5 --no-cpu-baseline --e2e-steps 3 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; l=d['config']['launch']
print('WPE=$w LSM_DEBUG=$d', 'step_ms', round(d['ms_per_step'],4), 'b2b', round(d['ms_per_step_back_to_back'],4), 'emit_ms', round(r['mean_launch_ms'],4), l['emit_block_threads'], l['emit_regs_per_thread'], l['emit_blocks_per_sm'])"
done; done
