#!/bin/bash
set -u
TAG=${1:-r3c}
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/${TAG}_pytest.log
for w in "cfg2 80 0 1" "cfg2 80 0 1" "cfg3 40 3 4" "cfg4 20 0 4" "cfg1 80 0 1"; do
timeout 300 python tools/tuning_sweep.py $w 2>&1 | grep -v Warn
done
