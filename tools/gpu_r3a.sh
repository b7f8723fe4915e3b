#!/bin/bash
set -u
TAG=${1:-r3a}
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "edge or world_graph or pair_value or golden_rollout" > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${TAG}_pytest.log
timeout 600 python tools/edge_bench.py > $O/${TAG}_edge_bench.jsonl 2>$O/${TAG}_edge_bench.err
python - <<PY
import json
for l in open('$O/${TAG}_edge_bench.jsonl'):
    d=json.loads(l); print(d['workload'], d['mode'], d['ms_per_step'], d['hbm_frac'], d['mean_degree'])
PY
timeout 300 python tools/tuning_sweep.py cfg3 40 3 4 2>&1 | grep -v Warn
timeout 300 python tools/tuning_sweep.py cfg2 60 0 1 2>&1 | grep -v Warn
