#!/usr/bin/env python
"""Host-side cost of one step: the bare C-ABI call vs B200GraphVecEnv.step (python + ctypes). The queue is kept short
(100 steps after a sync) so the host never blocks on the device. usage: tools/cpu_overhead_probe.py [workload]"""
import ctypes as C, os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'tests'))
import torch
import bench as B
from layered_safe_marl_b200 import B200GraphVecEnv
wl = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
args, flags, n, episode = B.build_args(wl)
env = B200GraphVecEnv(args, num_envs=min(n, 256), seed=1, binary_cfg=flags)     # tiny batch: the device is never the bottleneck
acts = torch.randint(0, 25, (env.n, env.N), device='cuda', dtype=torch.int32)
env.reset(episode)
for _ in range(20):
    env.step(acts, episode)
torch.cuda.synchronize()
K = 100
for rep in range(3):
    t0 = time.perf_counter()
    for _ in range(K):
        env.step(acts, episode)
    t1 = time.perf_counter(); torch.cuda.synchronize()
    ptr, st = C.c_void_p(acts.data_ptr()), env._stream()
    t2 = time.perf_counter()
    for _ in range(K):
        env.lib.lsm_step(env._h, ptr, None, episode, env.seed, 1, st)
    t3 = time.perf_counter(); torch.cuda.synchronize()
    print(f"{wl}: env.step {1e6*(t1-t0)/K:.1f} us/call on the host; bare lsm_step {1e6*(t3-t2)/K:.1f} us/call", flush=True)
