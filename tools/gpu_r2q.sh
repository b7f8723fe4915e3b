#!/bin/bash
set -u
TAG=${1:-r2q}
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/${TAG}_pytest.log
python tools/graph_probe.py cfg2 > $O/${TAG}_graph_probe.log 2>&1; tail -5 $O/${TAG}_graph_probe.log
python tools/graph_probe.py cfg1 >> $O/${TAG}_graph_probe.log 2>&1; tail -4 $O/${TAG}_graph_probe.log
python bench.py --no-extra-workloads > $O/${TAG}_bench_cfg2.json 2> $O/${TAG}_bench_cfg2.err; echo "bench rc=$?"; tail -3 $O/${TAG}_bench_cfg2.err
python - <<PY
import json
d=json.loads(open('$O/${TAG}_bench_cfg2.json').read().strip().splitlines()[-1])
print('step_ms', d['ms_per_step'], 'b2b', d['ms_per_step_back_to_back'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['value'], 'launch', d['config']['launch'])
PY
