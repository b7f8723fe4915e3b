#!/bin/bash
# timeline of the three kernels per workload: placement (late / LSM_DEBUG=32 up-front)
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in ${@:-cfg3 cfg4 cfg2}; do
  for d in 0 32; do LSM_DEBUG=$d python tools/timeline.py $w; done
  LSM_NO_PACKED=1 LSM_DEBUG=32 python tools/timeline.py $w
done
