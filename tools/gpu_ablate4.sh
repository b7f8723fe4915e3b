#!/bin/bash
# whole step + emit kernel alone against the number of resident emit blocks per SM (lab build, LSM_EMIT_BPS)
W=$1; shift
export LSM_LIB=$PWD/layered_safe_marl_b200/liblsm_b200_exp.so
for b in "$@"; do
  LSM_EMIT_BPS=$b python bench.py --workload $W --steps 40 --warmup 5 --no-cpu-baseline --no-extra-workloads --e2e-steps 2 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; dk=r.get('dominant_kernel') or {}
print('$W LSM_EMIT_BPS=$b', 'step_ms', round(d['ms_per_step'],4), 'b2b', round(d['ms_per_step_back_to_back'],4), 'emit_ms', round(dk.get('mean_launch_ms',0),4))"
done
