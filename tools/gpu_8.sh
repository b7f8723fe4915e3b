#!/bin/bash
# 8-GPU visit, the driver's own commands: reference arm, N=1 and N=8 default lines (K = 20 as the driver uses, then K = 200)
set -u
TAG=${1:-r02e}; N=${2:-8}
O=gpurun_out; mkdir -p $O
nproc
t0=$(date +%s)
python bench.py --gpus 1 --steps 20 --warmup 5 > $O/${TAG}_bench_1gpu_k20.json 2> $O/${TAG}_bench_1gpu_k20.err; echo "N=1 rc=$? $(( $(date +%s) - t0 )) s"
for K in 20 200; do
  t0=$(date +%s)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$((K/100)) bench.py --gpus $N --steps $K --warmup 5 > $O/${TAG}_bench_${N}gpu_k$K.json 2> $O/${TAG}_bench_${N}gpu_k$K.err; echo "N=$N K=$K rc=$? $(( $(date +%s) - t0 )) s"
done
t0=$(date +%s)
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --impl reference --gpus $N --steps 20 --warmup 5 > $O/${TAG}_ref_${N}gpu.json 2> $O/${TAG}_ref_${N}gpu.err; echo "ref N=$N rc=$? $(( $(date +%s) - t0 )) s"
python - <<PY
import json
for f in ("${TAG}_bench_1gpu_k20", "${TAG}_bench_${N}gpu_k20", "${TAG}_bench_${N}gpu_k200"):
    try:
        d=json.loads(open("$O/%s.json"%f).read().strip().splitlines()[-1])
        print(f, "value %.4g ms %.5f median %.5f b2b %.5f e2e %.4g (%.3f ms, %d thr)" % (d["value"], d["ms_per_step"], d["roofline"]["median_ms"], d["ms_per_step_back_to_back"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["host_side"]["host_threads"]), {k:(float("%.4g"%v["value"]), round(v["ms_per_step"],4)) for k,v in (d.get("workloads") or {}).items()})
    except Exception as e: print(f, "ERR", e)
r=json.loads(open("$O/${TAG}_ref_${N}gpu.json").read().strip().splitlines()[-1]); print("ref", r["value"], r["cpu_baseline"]["sample"])
PY
