#!/bin/bash
# full ncu capture of the step kernels. usage: tools/gpu_prof.sh <tag> <workload> [count]
TAG=$1; W=${2:-cfg2}; C=${3:-3}; O=gpurun_out
CMD="python bench.py --workload $W --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 3"
$CMD > $O/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:lsm_ -s 12 -c $C -f -o $O/${TAG}_prof_$W $CMD > $O/${TAG}_ncu2.log 2>&1
tail -3 $O/${TAG}_ncu2.log
