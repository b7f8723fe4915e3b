#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench lines of every workload, ncu launch list + full capture of cfg2.
# usage: tools/gpu_round.sh <tag>   (outputs under gpurun_out/<tag>_*)
set -u
TAG=${1:-run}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
python __graft_entry__.py smoke > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py > $O/${TAG}_bench_cfg2.json 2> $O/${TAG}_bench_cfg2.err; echo "bench rc=$?"
for w in cfg1 cfg3 cfg4; do
  python bench.py --workload $w --steps 50 --warmup 5 --no-cpu-baseline > $O/${TAG}_bench_$w.json 2> $O/${TAG}_bench_$w.err
done
python bench.py --impl reference --steps 20 --warmup 3 > $O/${TAG}_bench_ref.json 2>&1
CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 3"
$CMD > $O/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/${TAG}_launches.csv $CMD > $O/${TAG}_ncu1.log 2>&1
$CMD > $O/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:lsm_ -s 12 -c 6 -f -o $O/${TAG}_prof_cfg2 $CMD > $O/${TAG}_ncu2.log 2>&1
tail -3 $O/${TAG}_pytest.log; cat $O/${TAG}_smoke.log | tail -2
python - <<PY
import json
for w in ('cfg2','cfg1','cfg3','cfg4'):
    try:
        d=json.loads(open('$O/${TAG}_bench_%s.json'%w).read().strip().splitlines()[-1])
        r=d['roofline']
        print(w, 'step_ms', round(d['ms_per_step'],4), 'b2b', round(d['ms_per_step_back_to_back'],4), 'emit_ms', round(r['mean_launch_ms'],4), 'emit_frac', round(r['frac'],3), 'step_frac', round(r['whole_step']['frac'],3), 'e2e', d['e2e']['value'], 'cpu', (d.get('cpu_baseline') or {}).get('value'))
    except Exception as e: print(w, 'ERR', e)
PY
