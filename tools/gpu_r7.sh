#!/bin/bash
for rep in 1 2; do for p in late middle; do bash tools/gpu_dbg.sh cfg2 "0" LSM_PAIR=$p; done; done
for w in cfg3 cfg4; do for p in late middle; do bash tools/gpu_dbg.sh $w "0" LSM_PAIR=$p; done; done
