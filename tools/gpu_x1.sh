#!/bin/bash
# GPU visit for the obstacle extension: all GPU parity tests, the obstacle workload and the default line (short).
# usage: tools/gpu_x1.sh <tag>
TAG=${1:-x1}; O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 $O/${TAG}_pytest.log
timeout 300 python bench.py --workload cfg3_obst4 --steps 40 --warmup 5 --no-cpu-baseline --no-extra-workloads --e2e-steps 3 > $O/${TAG}_bench_cfg3_obst4.json 2> $O/${TAG}_bench_cfg3_obst4.err || tail -5 $O/${TAG}_bench_cfg3_obst4.err
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-extra-workloads --e2e-steps 5 > $O/${TAG}_bench_cfg2.json 2> $O/${TAG}_bench_cfg2.err || tail -5 $O/${TAG}_bench_cfg2.err
python - <<PY
import json
for w in ('cfg3_obst4', 'cfg2'):
    try:
        d=json.loads(open('$O/${TAG}_bench_%s.json'%w).read().strip().splitlines()[-1])
        print(w, 'value', d['value'], 'step_ms', round(d['ms_per_step'],4), 'b2b', round(d['ms_per_step_back_to_back'],4), 'frac', round(d['roofline']['frac'],3), 'e2e_ms', d['e2e'].get('ms_per_step'), d['config'].get('launch'))
    except Exception as e: print(w, 'ERR', e)
PY
