#!/bin/bash
# round-2 probe visit: parity tests, host-facing step sweep, CUDA-graph replay probe, launch-shape sweep
set -u
TAG=${1:-r2p}
O=gpurun_out; mkdir -p $O
nproc; lscpu | grep -E "Model name|Socket|NUMA node\(s\)|Thread" 
python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/${TAG}_pytest.log
python tools/e2e_probe.py cfg2 > $O/${TAG}_e2e_probe.log 2>&1; cat $O/${TAG}_e2e_probe.log
python tools/graph_probe.py cfg2 > $O/${TAG}_graph_probe.log 2>&1; tail -4 $O/${TAG}_graph_probe.log
python tools/tuning_sweep.py cfg2 40 > $O/${TAG}_sweep_cfg2.log 2>&1; tail -12 $O/${TAG}_sweep_cfg2.log
