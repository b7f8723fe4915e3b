#!/usr/bin/env python
"""Timing of the COO edge-list build (N2) on the step's adjacency at a benchmark size, against torch's nonzero path
(what TransformerConvNet.process_adj does).  usage: tools/edge_probe.py <workload>"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'tests'))
import torch
import bench as B
from layered_safe_marl_b200 import B200GraphVecEnv
w = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
args, flags, n, episode = B.build_args(w)
env = B200GraphVecEnv(args, num_envs=n, seed=1, binary_cfg=flags)
env.reset(episode)
for _ in range(5):
    env.step(torch.randint(0, 25, (n, env.N), device=env.device, dtype=torch.int32), episode)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=env.device)
def timeit(fn, K=20):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(K):
        flush.fill_(0.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]
ei, ea = env.edge_list()
nnz = ei.shape[1]
adj = env.adj
def torch_path():
    a = adj.view(-1, env.E, env.E)
    idx = a.nonzero(as_tuple=False)
    attr = a[idx[:, 0], idx[:, 1], idx[:, 2]]
    b = idx[:, 0] * env.E
    return torch.stack([b + idx[:, 1], b + idx[:, 2]]), attr
t_ours = timeit(lambda: env.edge_list())
try:
    t_torch = timeit(torch_path)
except torch.OutOfMemoryError:
    t_torch = float('nan')       # nonzero() materialises (nnz, 3) int64 indices: tens of GB at the dense 32-agent size
bytes_alg = adj.numel() * 4 + nnz * (16 + 4)
print(w, 'graphs', n * env.N, 'nnz', nnz, 'ours ms', round(t_ours, 4), 'torch nonzero path ms', round(t_torch, 4),
      'algorithmic GB/s (adj read once + edges written)', round(bytes_alg / t_ours / 1e6, 1))
