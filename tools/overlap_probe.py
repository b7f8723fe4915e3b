#!/usr/bin/env python
"""Do two half-size shards stepped on two CUDA streams overlap (agent kernel of one beside the emit kernel of the
other)?  usage: tools/overlap_probe.py <workload> [shards]"""
import os, sys, time
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'tests'))
import torch
import bench as B
from layered_safe_marl_b200 import B200GraphVecEnv

w = sys.argv[1]; S = int(sys.argv[2]) if len(sys.argv) > 2 else 2
args, flags, n_envs, episode = B.build_args(w)
K = 40
dev = torch.device('cuda:0')
def run(shards, stagger):
    envs = [B200GraphVecEnv(args, num_envs=n_envs // shards, seed=1234, binary_cfg=flags, env_id_base=k * (n_envs // shards)) for k in range(shards)]
    streams = [torch.cuda.Stream() for _ in range(shards)]
    acts = [torch.randint(0, 25, (K + 10, n_envs // shards, envs[0].N), device=dev, dtype=torch.int32) for _ in range(shards)]
    for e in envs: e.reset(episode)
    torch.cuda.synchronize()
    def step_all(t):
        for k, (e, s) in enumerate(zip(envs, streams)):
            with torch.cuda.stream(s):
                e.step(acts[k][t], episode)
    for t in range(10): step_all(t)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in range(K): step_all(10 + t)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / K
    for e in envs: e.close()
    return dt * 1e3
print(w, 'single shard   ms/step', round(run(1, False), 4))
for s in (2, 4):
    if s <= S: print(w, f'{s} shards/streams ms/step', round(run(s, False), 4))
