#!/usr/bin/env python
"""Overlap experiment (lab build): does a step get shorter when the agent kernel is held to ONE block per SM and the batch is
split into env ranges, so that range k's emit / pair kernels run on the free half of every SM beside range k+1's agent
kernel? Launch overhead is taken out with a CUDA graph captured around env.step (torch captures the fork / join).
usage: LSM_LIB=layered_safe_marl_b200/liblsm_b200_exp.so [LSM_AGENT_SMEM=120000] tools/overlap2_probe.py <workload> <chunks> [placement]"""
import os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'tests'))
import torch
import bench as B
from layered_safe_marl_b200 import B200GraphVecEnv

wl = sys.argv[1]; chunks = int(sys.argv[2]); placement = int(sys.argv[3]) if len(sys.argv) > 3 else -1
args, flags, n, episode = B.build_args(wl)
K, W = 100, 8
env = B200GraphVecEnv(args, num_envs=n, seed=1234, binary_cfg=flags, tuning=dict(chunks=chunks, pair_placement=placement, use_graph=0))
acts = torch.randint(0, 25, (K + W, n, env.N), device='cuda', dtype=torch.int32)
static = acts[0].clone()
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device='cuda')
env.reset(episode)
for t in range(W):
    env.step(acts[t], episode)
torch.cuda.synchronize()

def timed(fn):
    out = []
    for flushed in (True, False):
        st = [torch.cuda.Event(enable_timing=True) for _ in range(K)]; en = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
        for _ in range(8):
            flush.fill_(0.0)
        for t in range(K):
            if flushed:
                flush.fill_(0.0)
            st[t].record(); fn(t); en[t].record()
        torch.cuda.synchronize()
        per = np.array([a.elapsed_time(b) for a, b in zip(st, en)]) * 1e3
        out.append(float(per.mean()) if flushed else st[0].elapsed_time(en[-1]) / K * 1e3)
    return out

plain = timed(lambda t: env.step(acts[W + t], episode))
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    env.step(static, episode); env.step(static, episode)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        env.step(static, episode)
torch.cuda.synchronize()
graph = timed(lambda t: g.replay())
li = env.launch_info()
print(f"{wl} chunks={li['chunks']} placement={li['pair_placement']} agent_smem={os.environ.get('LSM_AGENT_SMEM', '-')}: "
      f"plain flushed {plain[0]:.1f} b2b {plain[1]:.1f} us | graph flushed {graph[0]:.1f} b2b {graph[1]:.1f} us", flush=True)
