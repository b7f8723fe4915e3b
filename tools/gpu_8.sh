#!/bin/bash
# 8-GPU visit: default bench line with plain launches and with graph replay (scaling jitter), K as the driver uses
set -u
TAG=${1:-r2s}; N=${2:-8}
O=gpurun_out; mkdir -p $O
nproc
for g in 0 1; do
  for K in 20 200; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$g bench.py --gpus $N --steps $K --warmup 5 --no-cpu-baseline --no-extra-workloads --use-graph $g > $O/${TAG}_bench_${N}gpu_g${g}_k$K.json 2> $O/${TAG}_bench_${N}gpu_g${g}_k$K.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/${TAG}_bench_${N}gpu_g${g}_k$K.json").read().strip().splitlines()[-1])
    print("N=$N graph=$g K=$K value", d["value"], "ms", d["ms_per_step"], "median", d["roofline"].get("median_ms"), "b2b", d["ms_per_step_back_to_back"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["host_side"]["host_threads"])
except Exception as e: print("ERR", e)
PY
  done
done
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-workloads --use-graph 0 > $O/${TAG}_bench_1gpu_g0.json 2>/dev/null
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-workloads --use-graph 1 > $O/${TAG}_bench_1gpu_g1.json 2>/dev/null
python - <<PY
import json
for g in (0,1):
    d=json.loads(open("$O/${TAG}_bench_1gpu_g%d.json"%g).read().strip().splitlines()[-1])
    print("N=1 graph=%d"%g, d["value"], "ms", d["ms_per_step"], "median", d["roofline"].get("median_ms"), "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
PY
