#!/bin/bash
# pure-copy time of the emit kernel (LSM_DEBUG=256: the bulk copies alone) against the number of resident emit blocks per SM
W=$1; shift
export LSM_LIB=$PWD/layered_safe_marl_b200/liblsm_b200_exp.so
for spec in "$@"; do
  d=${spec%%:*}; b=${spec##*:}
  LSM_DEBUG=$d LSM_EMIT_BPS=$b python bench.py --workload $W --steps 20 --warmup 5 --no-cpu-baseline --no-extra-workloads --e2e-steps 2 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; dk=r.get('dominant_kernel') or {}
print('$W LSM_DEBUG=$d LSM_EMIT_BPS=$b', 'emit_ms', round(dk.get('mean_launch_ms',0),4))"
done
