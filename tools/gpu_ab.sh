#!/bin/bash
# A/B timing of library builds: usage tools/gpu_ab.sh <workload> <lib1.so> <lib2.so> ...   ("-" = the product library)
W=$1; shift
for lib in "$@"; do
  if [ "$lib" = "-" ]; then unset LSM_LIB; else export LSM_LIB=/root/repo/layered_safe_marl_b200/$lib; fi
  for rep in 1 2; do
  python bench.py --workload $W --steps 150 --warmup 10 --no-cpu-baseline --e2e-steps 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); t=d['timeline_us']
print('$lib', '$W', 'flushed %.2f us  b2b %.2f us  emit-alone %.2f us' % (d['ms_per_step']*1e3, d['ms_per_step_back_to_back']*1e3, d['roofline']['mean_launch_ms']*1e3), 'agent_end', t['agent_end'], 'emit', t['emit_start'], t['emit_end'], 'pair', t['pair_start'], t['pair_end'])"
  done
done
