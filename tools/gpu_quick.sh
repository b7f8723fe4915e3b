#!/bin/bash
# quick GPU visit: parity tests + short bench lines. usage: tools/gpu_quick.sh <tag> [workloads...]
TAG=${1:-q}; shift
WL=${@:-cfg2 cfg1 cfg3 cfg4}
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $O/${TAG}_pytest.log
for w in $WL; do
  timeout 300 python bench.py --workload $w --steps 100 --warmup 10 --no-cpu-baseline --e2e-steps 5 > $O/${TAG}_bench_$w.json 2> $O/${TAG}_bench_$w.err || tail -5 $O/${TAG}_bench_$w.err
done
python - <<PY
import json
for w in "$WL".split():
    try:
        d=json.loads(open('$O/${TAG}_bench_%s.json'%w).read().strip().splitlines()[-1])
        r=d['roofline']
        print(w, 'step_ms', round(d['ms_per_step'],4), 'b2b', round(d['ms_per_step_back_to_back'],4), 'emit_ms', round(r['mean_launch_ms'],4), 'emit_frac', round(r['frac'],3), 'step_frac', round(r['whole_step']['frac'],3), 'e2e', round(d['e2e']['value']/1e6,2), d['config']['launch'])
    except Exception as e: print(w, 'ERR', e)
PY
