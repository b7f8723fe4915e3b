"""How fast can the step's output tensors be written at all? torch fill_ of adj / node_obs after an L2 flush."""
import torch, json
dev = torch.device('cuda:0')
n, N, E, F = 4096, 8, 24, 10
adj = torch.empty((n, N, E, E), dtype=torch.float32, device=dev)
node = torch.empty((n, N, E, F), dtype=torch.float32, device=dev)
both = torch.empty(adj.numel() + node.numel(), dtype=torch.float32, device=dev)
fl = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
empty = torch.empty(1, device=dev)
def timed(fn, flush=True, iters=30):
    ts = []
    for _ in range(iters):
        if flush: fl.fill_(0.0)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e) * 1e3)
    ts = ts[5:]
    return sum(ts) / len(ts)
out = {}
out['empty_fill_us'] = timed(lambda: empty.fill_(0.0))
out['adj_fill_us'] = timed(lambda: adj.fill_(1.0))
out['node_fill_us'] = timed(lambda: node.fill_(1.0))
out['both_one_launch_us'] = timed(lambda: both.fill_(1.0))
out['adj_then_node_us'] = timed(lambda: (adj.fill_(1.0), node.fill_(1.0)))
out['both_noflush_us'] = timed(lambda: both.fill_(1.0), flush=False)
src = torch.randn_like(both)
out['copy_both_us'] = timed(lambda: both.copy_(src))
print(json.dumps(out, indent=1))
