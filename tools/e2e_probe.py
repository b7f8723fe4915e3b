#!/usr/bin/env python
"""Where the host-facing step (numpy_outputs=True) spends its time: kernels, D2H waits, host expansion.
usage: tools/e2e_probe.py [workload] [host_threads]"""
import os, sys, time
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'tests'))
import torch
import bench as B
from layered_safe_marl_b200 import B200GraphVecEnv

wl = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
args, flags, n, episode = B.build_args(wl)
for threads, cached in ([(int(sys.argv[2]), False)] if len(sys.argv) > 2 else [(4, False), (8, False), (16, False), (8, True), (16, True)]):
    env = B200GraphVecEnv(args, num_envs=n, seed=1, binary_cfg=flags, numpy_outputs=True, host_threads=threads)
    env.host_cached_stores = cached
    env.reset(episode)
    rng = np.random.default_rng(0)
    K = 30
    onehot = torch.from_numpy(np.eye(25, dtype=np.float32)[rng.integers(0, 25, (K + 3, n, env.N))]).pin_memory()
    for t in range(3):
        env.step(onehot[t].numpy(), episode)
    env.host_profile = {}
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for t in range(K):
        env.step(onehot[3 + t].numpy(), episode)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / K
    p = env.host_profile
    print(f"{wl} threads={threads} cached={cached}: {dt*1e3:.3f} ms/step; wait_chunk {p['wait_chunk']/K*1e3:.3f} expand {p['expand']/K*1e3:.3f} "
          f"final_sync {p['final_sync']/K*1e3:.3f} other {(dt - (p['wait_chunk']+p['expand']+p['final_sync'])/K)*1e3:.3f}", flush=True)
    env.close()
