#!/bin/bash
set -u
TAG=${1:-r2w}
O=gpurun_out; mkdir -p $O
for spec in "cfg4sparse edges" "cfg4sparse dense" "cfg2 edges" "cfg2 dense+edges" "cfg4sparse edges 1"; do
  set -- $spec
  name=${1}_$(echo $2 | tr '+' 'p')${3:+_c$3}
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_launch_$name.csv python tools/edge_prof.py $@ > $O/${TAG}_$name.log 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("$O/${TAG}_launch_$name.csv")) if len(r)>10 and r[0].isdigit()]
# last step's kernels: take the tail
import collections
tail=rows[-16:]
print("== $name")
for r in tail: print("  ", r[4][:60], r[7], r[8], r[-1])
PY
done
