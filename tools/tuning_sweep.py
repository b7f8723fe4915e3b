#!/usr/bin/env python
"""Sweep the product's launch-shape knobs (lsm_tuning: pair placement x env ranges) for a workload.
usage: tools/tuning_sweep.py <workload> [steps] [placements, e.g. 0,4] [chunks, e.g. 1,4]"""
import os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'tests'))
import torch
import bench as B
from layered_safe_marl_b200 import B200GraphVecEnv
wl = sys.argv[1]; K = int(sys.argv[2]) if len(sys.argv) > 2 else 60
args, flags, n, episode = B.build_args(wl)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device='cuda')
acts = torch.randint(0, 25, (K + 8, n, args.num_agents), device='cuda', dtype=torch.int32)
PL = [int(v) for v in sys.argv[3].split(',')] if len(sys.argv) > 3 else [0, 3, 2, 4]
CH = [int(v) for v in sys.argv[4].split(',')] if len(sys.argv) > 4 else [1, 2, 4, 8]
for placement in PL:
    for chunks in CH:
        env = B200GraphVecEnv(args, num_envs=n, seed=1234, binary_cfg=flags, tuning=dict(pair_placement=placement, chunks=chunks))
        env.reset(episode)
        for t in range(8):
            env.step(acts[t], episode)
        torch.cuda.synchronize()
        st = [torch.cuda.Event(enable_timing=True) for _ in range(K)]; en = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
        for t in range(K):
            flush.fill_(0.0); st[t].record(); env.step(acts[8 + t], episode); en[t].record()
        torch.cuda.synchronize()
        ms = float(np.mean([a.elapsed_time(b) for a, b in zip(st, en)]))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for t in range(K):
            env.step(acts[8 + t], episode)
        e1.record(); torch.cuda.synchronize()
        print(f"{wl} placement={placement} chunks={env.launch_info()['chunks']}: flushed {ms*1e3:.1f} us  b2b {e0.elapsed_time(e1)/K*1e3:.1f} us", flush=True)
        env.close(); del env
        torch.cuda.empty_cache()
