#!/bin/bash
# r7 visit: parity + bench of the late-pair pipeline, then emit-occupancy / placement variants
TAG=${1:-r7a}
bash tools/gpu_quick.sh $TAG cfg2 cfg1 cfg3 cfg4
echo "--- variants cfg2"
for b in 5 6 7 8; do bash tools/gpu_dbg.sh cfg2 "0" LSM_EMIT_BPS=$b; done
bash tools/gpu_dbg.sh cfg2 "32"
echo "--- variants cfg4"
bash tools/gpu_dbg.sh cfg4 "32"
for b in 2; do bash tools/gpu_dbg.sh cfg4 "0" LSM_EMIT_BPS=$b; done
echo "--- variants cfg3"
bash tools/gpu_dbg.sh cfg3 "32"
for b in 3 4; do bash tools/gpu_dbg.sh cfg3 "0" LSM_EMIT_BPS=$b; done
