#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for w in cfg3 cfg4 cfg2; do bash tools/gpu_dbg.sh $w "0 8192"; LSM_PAIR=front python tools/timeline.py $w | grep pair_end; done
