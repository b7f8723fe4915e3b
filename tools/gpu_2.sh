#!/bin/bash
set -u
TAG=${1:-r2u}; N=${2:-2}
O=gpurun_out; mkdir -p $O
for rep in 1 2 3; do
  LSM_BENCH_DUMP=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$rep bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-extra-workloads --e2e-steps 3 > $O/${TAG}_n${N}_$rep.json 2> $O/${TAG}_n${N}_$rep.err
  grep "per-step" $O/${TAG}_n${N}_$rep.err
  python - <<PY
import json
d=json.loads(open("$O/${TAG}_n${N}_$rep.json").read().strip().splitlines()[-1])
print("N=$N rep $rep value", d["value"], "ms", d["ms_per_step"], "median", d["roofline"].get("median_ms"))
PY
done
LSM_BENCH_DUMP=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-workloads --e2e-steps 3 2>&1 >$O/${TAG}_n1.json | grep per-step
python - <<PY
import json
d=json.loads(open("$O/${TAG}_n1.json").read().strip().splitlines()[-1])
print("N=1 value", d["value"], "ms", d["ms_per_step"], "median", d["roofline"].get("median_ms"))
PY
