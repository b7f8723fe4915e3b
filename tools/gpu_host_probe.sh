#!/bin/bash
# host-side facts the e2e path depends on: cores, NUMA, PCIe D2H bandwidth, host streaming-store bandwidth
O=gpurun_out; mkdir -p $O
{
nproc; lscpu | head -30; cat /sys/devices/system/node/node*/cpulist 2>/dev/null; nvidia-smi topo -m
for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/class 2>/dev/null)" = "0x030200" ]; then echo "$d numa=$(cat $d/numa_node)"; fi; done
which numactl; free -g | head -2
./tools/host_probe 108 64
python - <<'PY'
import torch, time
for mb in (8, 42, 108):
    n = mb << 20
    d = torch.empty(n, dtype=torch.uint8, device='cuda'); h = torch.empty(n, dtype=torch.uint8).pin_memory()
    for _ in range(3): h.copy_(d, non_blocking=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): h.copy_(d, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    print(f"D2H pinned {mb} MB: {dt*1e3:.3f} ms  {n/dt/1e9:.1f} GB/s")
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    print(f"H2D pinned {mb} MB: {dt*1e3:.3f} ms  {n/dt/1e9:.1f} GB/s")
PY
} > $O/host_probe.log 2>&1
tail -40 $O/host_probe.log
