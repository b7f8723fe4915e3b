"""Pure-write and copy bandwidth probes on this GPU (context for the write-dominated step kernel)."""
import torch, json
dev = torch.device('cuda:0')
def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e-3
out = {}
for mb in (110, 512, 2048, 8192):
    n = mb * 1024 * 1024 // 4
    a = torch.empty(n, dtype=torch.float32, device=dev)
    t = timeit(lambda: a.fill_(1.0))
    out[f'fill_{mb}MB_GBps'] = mb * 1.048576e6 / t / 1e9
    if mb <= 2048:
        b = torch.empty(n, dtype=torch.float32, device=dev)
        t = timeit(lambda: b.copy_(a))
        out[f'copy_{mb}MB_rw_GBps'] = 2 * mb * 1.048576e6 / t / 1e9
        del b
    del a
# fill with a flush in between (cold L2 with dirty lines), like bench.py's timed step
fl = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
a = torch.empty(110 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
ts = []
for _ in range(20):
    fl.fill_(0.0)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); a.fill_(1.0); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e) * 1e-3)
out['fill_110MB_after_flush_GBps'] = 110 * 1.048576e6 / (sum(ts[3:]) / len(ts[3:])) / 1e9
out['fill_110MB_after_flush_us'] = 1e6 * sum(ts[3:]) / len(ts[3:])
print(json.dumps(out, indent=1))
