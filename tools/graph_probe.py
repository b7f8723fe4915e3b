#!/usr/bin/env python
"""lsm_tuning.use_graph: plain launches vs the library's CUDA-graph replay of one env.step (device time per step, flushed
and back to back, and host time per step). usage: tools/graph_probe.py [workload]"""
import os, sys, time
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'tests'))
import torch
import bench as B
from layered_safe_marl_b200 import B200GraphVecEnv

wl = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
args, flags, n, episode = B.build_args(wl)
K, W = 150, 10
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device='cuda')
for use_graph in (0, 1, 0, 1):
    env = B200GraphVecEnv(args, num_envs=n, seed=1234, binary_cfg=flags, tuning=dict(use_graph=use_graph))
    acts = torch.randint(0, 25, (K + W, n, env.N), device='cuda', dtype=torch.int32)
    static = acts[0].clone()
    env.reset(episode)
    for t in range(W):
        static.copy_(acts[t]); env.step(static, episode)
    torch.cuda.synchronize()
    res = []
    for flushed in (True, False):
        st = [torch.cuda.Event(enable_timing=True) for _ in range(K)]; en = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
        t_step = 0.0
        for t in range(K):
            static.copy_(acts[W + t])
            if flushed:
                flush.fill_(0.0)
            st[t].record()
            c0 = time.perf_counter(); env.step(static, episode); t_step += time.perf_counter() - c0
            en[t].record()
        torch.cuda.synchronize()
        per = np.array([a.elapsed_time(b) for a, b in zip(st, en)]) * 1e3
        dev = float(per.mean()) if flushed else st[0].elapsed_time(en[-1]) / K * 1e3
        res.append((dev, float(np.median(per)), float(per.max()), t_step / K * 1e6))
    li = env.launch_info()
    print(f"{wl} use_graph={use_graph}: flushed mean {res[0][0]:.2f} median {res[0][1]:.2f} max {res[0][2]:.1f} us (host {res[0][3]:.1f} us/step) | "
          f"b2b {res[1][0]:.2f} us (host {res[1][3]:.1f}) | replays {li['graph_replays']} captures {li['graph_captures']}", flush=True)
    env.close()
