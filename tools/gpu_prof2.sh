#!/bin/bash
O=gpurun_out
python tools/fill_probe.py > $O/fillprobe_plain.log 2>&1 &&
ncu --set full --clock-control none -k regex:vectorized_elementwise -s 40 -c 4 -f -o $O/prof_fill python tools/fill_probe.py > $O/prof_fill.log 2>&1
CMD="python bench.py --workload cfg2 --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 3"
LSM_DEBUG=288 $CMD > $O/so_plain.log 2>&1 &&
LSM_DEBUG=288 ncu --set full --clock-control none -k regex:lsm_emit -s 10 -c 2 -f -o $O/prof_storesonly $CMD > $O/prof_so.log 2>&1
$CMD > $O/full_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:lsm_ -s 12 -c 4 -f -o $O/v5a_prof_cfg2 $CMD > $O/prof_full.log 2>&1
ls -la $O/*.ncu-rep | tail -4
