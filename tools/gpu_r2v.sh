#!/bin/bash
set -u
TAG=${1:-r2v}
O=gpurun_out; mkdir -p $O
python -m pytest tests/test_cuda_parity.py -m gpu -x -q -k "pair_value_variants or pair_tail or graph_replay" > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${TAG}_pytest.log
python tools/tuning_sweep.py cfg2 60 0,4 1 > $O/${TAG}_sweep.log 2>&1
python tools/tuning_sweep.py cfg2 60 0,4 1 >> $O/${TAG}_sweep.log 2>&1
python tools/tuning_sweep.py cfg3 40 3,4,0 4,1 >> $O/${TAG}_sweep.log 2>&1
python tools/tuning_sweep.py cfg4 20 0,4 4 >> $O/${TAG}_sweep.log 2>&1
python tools/tuning_sweep.py cfg1 60 0,4 1 >> $O/${TAG}_sweep.log 2>&1
grep -v Warn $O/${TAG}_sweep.log
