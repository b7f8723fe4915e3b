#!/bin/bash
set -u
TAG=${1:-r2t}
O=gpurun_out; mkdir -p $O
export LSM_LIB=layered_safe_marl_b200/liblsm_b200_exp.so
L=$O/${TAG}_overlap.log; : > $L
for spec in "cfg2 1 0 0" "cfg2 2 0 0" "cfg2 2 0 120000" "cfg2 2 3 120000" "cfg2 3 0 120000" "cfg2 4 0 120000" "cfg2 4 3 120000" \
            "cfg3 4 3 0" "cfg3 4 3 120000" "cfg3 8 3 120000" "cfg3 4 0 120000" "cfg3 8 0 120000" \
            "cfg4 4 0 0" "cfg4 4 0 120000" "cfg4 8 0 120000"; do
  set -- $spec
  LSM_AGENT_SMEM=$4 timeout 120 python tools/overlap2_probe.py $1 $2 $3 >> $L 2>&1 || echo "FAILED $spec" >> $L
done
grep -v Warning $L | tail -20
