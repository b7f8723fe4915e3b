#!/usr/bin/env python
"""Host-facing step (numpy_outputs=True, one lsm_step_host call per step): ms per step for host thread / env-range counts.
usage: tools/e2e_probe.py [workload] [threads,chunks ...]"""
import os, sys, time
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'tests'))
import torch
import bench as B
from layered_safe_marl_b200 import B200GraphVecEnv

wl = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
args, flags, n, episode = B.build_args(wl)
combos = [tuple(int(v) for v in a.split(',')) for a in sys.argv[2:]] or [(8, 4), (8, 8), (16, 4), (16, 8), (16, 16), (24, 8), (32, 8), (32, 16)]
print('cpus', len(os.sched_getaffinity(0)), flush=True)
for threads, chunks in combos:
    for cached in (False, True):
        env = B200GraphVecEnv(args, num_envs=n, seed=1, binary_cfg=flags, numpy_outputs=True, host_threads=threads, host_chunks=chunks,
                              host_cached_stores=cached)
        env.reset(episode)
        rng = np.random.default_rng(0)
        K = 30
        onehot = torch.from_numpy(np.eye(25, dtype=np.float32)[rng.integers(0, 25, (K + 3, n, env.N))]).pin_memory()
        for t in range(3):
            env.step(onehot[t].numpy(), episode)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for t in range(K):
            env.step(onehot[3 + t].numpy(), episode)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / K
        print(f"{wl} threads={threads} chunks={chunks} cached={cached}: {dt*1e3:.3f} ms/step", flush=True)
        env.close()
