#!/bin/bash
# ncu --set full of the agent + pair kernels of one workload (source-level), launch list first
set -u
TAG=${1:-r2p3}; W=${2:-cfg3}
O=gpurun_out; mkdir -p $O
CMD="python bench.py --workload $W --steps 4 --warmup 3 --no-cpu-baseline --no-extra-workloads --e2e-steps 3"
$CMD > $O/${TAG}_plain.log 2>&1 && echo plain ok
ncu --set full --clock-control none --import-source on -k regex:lsm_agent -s 8 -c 2 -f -o $O/${TAG}_agent_$W $CMD > $O/${TAG}_ncu_agent.log 2>&1; echo "ncu agent rc=$?"
ncu --set full --clock-control none --import-source on -k regex:lsm_pair -s 8 -c 2 -f -o $O/${TAG}_pair_$W $CMD > $O/${TAG}_ncu_pair.log 2>&1; echo "ncu pair rc=$?"
ls -la $O/${TAG}_*
