#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for w in cfg3 cfg4; do for c in 1 2 3 4 6 8; do bash tools/gpu_dbg.sh $w "0" LSM_CHUNKS=$c; done; done
for c in 1 2; do bash tools/gpu_dbg.sh cfg2 "0" LSM_CHUNKS=$c; done
