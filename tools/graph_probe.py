#!/usr/bin/env python
"""Does replaying one env.step from a captured CUDA graph beat three plain launches? (probe for lsm_tuning.use_graph)
usage: tools/graph_probe.py [workload]"""
import os, sys, time
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'tests'))
import torch
import bench as B
from layered_safe_marl_b200 import B200GraphVecEnv

wl = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
args, flags, n, episode = B.build_args(wl)
env = B200GraphVecEnv(args, num_envs=n, seed=1234, binary_cfg=flags)
K, W = 150, 10
acts = torch.randint(0, 25, (K + W, n, env.N), device='cuda', dtype=torch.int32)
static = acts[0].clone()
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device='cuda')
env.reset(episode)
for t in range(W):
    env.step(acts[t], episode)
torch.cuda.synchronize()

def timed(fn, flushed):
    st = [torch.cuda.Event(enable_timing=True) for _ in range(K)]; en = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    t0 = time.perf_counter()
    for t in range(K):
        if flushed:
            flush.fill_(0.0)
        st[t].record(); fn(t); en[t].record()
    t_cpu = (time.perf_counter() - t0) / K
    torch.cuda.synchronize()
    if flushed:
        return float(np.mean([a.elapsed_time(b) for a, b in zip(st, en)])) * 1e3, t_cpu * 1e6
    return st[0].elapsed_time(en[-1]) / K * 1e3, t_cpu * 1e6

def plain(t):
    env.step(acts[W + t], episode)
print(wl, 'plain   flushed %.2f us (cpu %.1f us/step)' % timed(plain, True), ' b2b %.2f us (cpu %.1f)' % timed(plain, False))

s = torch.cuda.Stream()
with torch.cuda.stream(s):
    env.step(static, episode)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        env.step(static, episode)
torch.cuda.synchronize()
def replay(t):
    static.copy_(acts[W + t])
    g.replay()
def replay_only(t):
    g.replay()
print(wl, 'graph   flushed %.2f us (cpu %.1f us/step)' % timed(replay_only, True), ' b2b %.2f us (cpu %.1f)' % timed(replay_only, False))
print(wl, 'graph+copy flushed %.2f us (cpu %.1f us/step)' % timed(replay, True), ' b2b %.2f us (cpu %.1f)' % timed(replay, False))
