#!/bin/bash
# emit-kernel ablations on the lab build (LSM_DEBUG bits: 4 no node rows, 8 no adjacency stores, 64 launch floor, 128 + record
# load, 256 bulk copies alone, 1024 / 2048 the same bytes as plain 16-byte stores). usage: tools/gpu_ablate2.sh <workload> <bits...>
W=$1; shift
export LSM_LIB=$PWD/layered_safe_marl_b200/liblsm_b200_exp.so
for d in "$@"; do
  LSM_DEBUG=$d python bench.py --workload $W --steps 30 --warmup 5 --no-cpu-baseline --no-extra-workloads --e2e-steps 2 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; dk=r.get('dominant_kernel') or {}
print('$W LSM_DEBUG=$d', 'step_ms', round(d['ms_per_step'],4), 'emit_ms', round(dk.get('mean_launch_ms',0),4))"
done
