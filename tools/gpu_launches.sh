#!/bin/bash
# per-kernel durations of a short bench run (ncu launch list). usage: tools/gpu_launches.sh <tag> <workload>
TAG=$1; W=${2:-cfg2}; O=gpurun_out
CMD="python bench.py --workload $W --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 3"
$CMD > $O/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/${TAG}_launches_$W.csv $CMD > $O/${TAG}_ncu1.log 2>&1
python - <<PY
import csv
rows=list(csv.reader(l for l in open('$O/${TAG}_launches_$W.csv') if l.startswith('"')))
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); gi=h.index('Grid Size'); bi=h.index('Block Size')
for r in rows[1:]:
    if 'lsm' in r[ki]: print(r[ki][:60], r[gi], r[bi], r[vi])
PY
