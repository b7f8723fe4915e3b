#!/usr/bin/env python
"""N2: step time with the fused COO edge output, dense adjacency on / off, against the plain dense step.
usage: tools/edge_bench.py [workload ...]   (cfg2 cfg4 cfg4sparse)
Prints one JSON line per (workload, mode): ms per step (flushed), bytes written, mean degree."""
import json, os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'tests'))
import torch
import _golden as G
import bench as B
from layered_safe_marl_b200 import B200GraphVecEnv

EXTRA = {   # sparse variants: same agents, a world large enough that most pairs are beyond the 4.0 radius
    'cfg4sparse': (dict(dynamics_type='double_integrator', num_agents=32, num_landmarks=2, use_safety_filter=True,
                        world_size=40, episode_length=250), {}, 8192, 6249),
    'cfg2sparse': (dict(dynamics_type='double_integrator', num_agents=8, num_landmarks=2, use_safety_filter=True,
                        world_size=20, episode_length=250), {}, 4096, 6249),
}
K, W = 50, 8
peak, _ = B.measured_peak()
for wl in (sys.argv[1:] or ['cfg2', 'cfg2sparse', 'cfg4', 'cfg4sparse']):
    if wl in EXTRA:
        kw, flags, n, episode = EXTRA[wl]
        args, flags = G.default_args(**kw), G.BinaryFlags(flags)
    else:
        args, flags, n, episode = B.build_args(wl)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device='cuda')
    for mode in ('dense', 'dense+edges', 'edges'):
        env = B200GraphVecEnv(args, num_envs=n, seed=1234, binary_cfg=flags)
        if mode != 'dense':
            env.enable_edge_output(dense_adj=(mode == 'dense+edges'))
        gen = torch.Generator(device='cuda'); gen.manual_seed(1)
        acts = torch.randint(0, 25, (K + W, n, env.N), generator=gen, device='cuda', dtype=torch.int32)
        env.reset(episode)
        for t in range(W):
            env.step(acts[t], episode)
        torch.cuda.synchronize()
        st = [torch.cuda.Event(enable_timing=True) for _ in range(K)]; en = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
        for t in range(K):
            flush.fill_(0.0)
            st[t].record(); env.step(acts[W + t], episode); en[t].record()
        torch.cuda.synchronize()
        ms = float(np.mean([a.elapsed_time(b) for a, b in zip(st, en)]))
        N, E, F, D = env.N, env.E, env.F, env.D
        small = 4 * N * (E * F + D + 2) + 77 * N
        nnz = int(env.edge_offsets[-1].item()) if mode != 'dense' else None
        per_env = small + (4 * N * E * E if mode != 'edges' else 0) + (20 * nnz / n + 12 * N if nnz is not None else 0)
        line = {"workload": wl, "mode": mode, "envs": n, "N": N, "E": E, "ms_per_step": round(ms, 4),
                "agent_steps_per_s": round(n * N / (ms / 1e3)), "algorithmic_bytes_per_env_step": round(per_env),
                "hbm_frac": round(per_env * n / (ms / 1e3) / 1e9 / peak, 3),
                "mean_degree": None if nnz is None else round(nnz / (n * N * E), 2), "launch": env.launch_info()['launches_per_step']}
        print(json.dumps(line), flush=True)
        env.close(); del env, acts
        torch.cuda.empty_cache()
