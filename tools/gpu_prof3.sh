#!/bin/bash
# ncu --set full of the agent kernel of cfg3 and cfg2 (source-level)
set -u
TAG=${1:-r3e}
O=gpurun_out; mkdir -p $O
for W in cfg3 cfg2; do
CMD="python bench.py --workload $W --steps 4 --warmup 3 --no-cpu-baseline --no-extra-workloads --e2e-steps 3"
$CMD > $O/${TAG}_plain_$W.log 2>&1 && echo plain ok
ncu --set full --clock-control none --import-source on -k regex:lsm_agent -s 8 -c 1 -f -o $O/${TAG}_agent_$W $CMD > $O/${TAG}_ncu_agent_$W.log 2>&1; echo "ncu agent $W rc=$?"
done
