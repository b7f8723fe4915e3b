#!/bin/bash
# emit-kernel ablations (LSM_DEBUG bits: 4 no node rows, 8 no adjacency stores, 32 no next-step pair values)
W=${1:-cfg2}
for d in 0 32 36 40 44; do
  LSM_DEBUG=$d python bench.py --workload $W --steps 60 --warmup 5 --no-cpu-baseline --e2e-steps 3 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('LSM_DEBUG=$d', 'step_ms', round(d['ms_per_step'],4), 'b2b', round(d['ms_per_step_back_to_back'],4), 'emit_ms', round(r['mean_launch_ms'],4))"
done
