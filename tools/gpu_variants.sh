#!/bin/bash
# A/B of library variants built with _build.build(extra_flags=..., out_path=liblsm_b200_<v>.so): same workload, same box.
# usage: tools/gpu_variants.sh <workload> <variant> [<variant> ...]     ("base" = the product library)
W=$1; shift
for v in "$@"; do
  if [ "$v" = base ]; then unset LSM_LIB; else export LSM_LIB=$PWD/layered_safe_marl_b200/liblsm_b200_$v.so; fi
  python bench.py --workload $W --steps 60 --warmup 6 --no-cpu-baseline --no-extra-workloads --e2e-steps 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; dk=r.get('dominant_kernel') or {}; li=d['config']['launch']
print('$W $v', 'step_ms', round(d['ms_per_step'],4), 'b2b', round(d['ms_per_step_back_to_back'],4), 'emit_ms', round(dk.get('mean_launch_ms',0),4), 'emit bps/threads/smem', li['emit_blocks_per_sm'], li['emit_block_threads'], li['emit_smem_bytes_per_block'], 'tl', d.get('timeline_us'))"
done
