#!/bin/bash
# One GPU-box visit for the round-2 evidence: parity tests, smoke, default bench line (+ reference arm), the other
# workloads as headline lines, the ncu launch list and one --set full capture of the three kernels of cfg2.
# usage: tools/gpu_round2.sh <tag>   (outputs under gpurun_out/<tag>_*)
set -u
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
python __graft_entry__.py smoke > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py > $O/${TAG}_bench_cfg2.json 2> $O/${TAG}_bench_cfg2.err; echo "bench rc=$?"
python bench.py --impl reference --steps 20 --warmup 3 > $O/${TAG}_bench_reference_arm.json 2>&1
for w in cfg1 cfg3 cfg3_obst4 cfg4; do
  python bench.py --workload $w --steps 60 --warmup 5 --no-cpu-baseline --no-extra-workloads --e2e-steps 5 > $O/${TAG}_bench_$w.json 2> $O/${TAG}_bench_$w.err
done
python tools/edge_bench.py > $O/${TAG}_edge_bench.jsonl 2> $O/${TAG}_edge_bench.err
python tools/e2e_probe.py cfg2 8,8 16,8 > $O/${TAG}_e2e_probe.log 2>&1
CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra-workloads --e2e-steps 3"
$CMD > $O/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/${TAG}_launches_cfg2.csv $CMD > $O/${TAG}_ncu1.log 2>&1
$CMD > $O/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:lsm_ -s 12 -c 6 -f -o $O/${TAG}_prof_cfg2 $CMD > $O/${TAG}_ncu2.log 2>&1
tail -3 $O/${TAG}_pytest.log; tail -2 $O/${TAG}_smoke.log
python - <<PY
import json
for w in ("cfg2","cfg1","cfg3","cfg3_obst4","cfg4"):
    try:
        d=json.loads(open('$O/${TAG}_bench_%s.json'%w).read().strip().splitlines()[-1])
        r=d['roofline']; dk=r.get('dominant_kernel') or {}
        print(w, 'step_ms', round(d['ms_per_step'],4), 'b2b', round(d['ms_per_step_back_to_back'],4), 'step_frac', round(r['frac'],3), 'emit_ms', round(dk.get('mean_launch_ms',0),4), 'emit_frac', round(dk.get('frac',0),3), 'e2e_ms', round(d['e2e']['ms_per_step'],3), 'cpu', (d.get('cpu_baseline') or {}).get('value'))
    except Exception as e: print(w, 'ERR', e)
PY
