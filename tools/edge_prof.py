#!/usr/bin/env python
"""A few steps of one edge-output mode (for an ncu launch list). usage: tools/edge_prof.py <cfg2|cfg2sparse|cfg4|cfg4sparse> <dense|dense+edges|edges> [chunks]"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'tests'))
import torch
import _golden as G
import bench as B
from layered_safe_marl_b200 import B200GraphVecEnv
wl, mode = sys.argv[1], sys.argv[2]
tuning = dict(chunks=int(sys.argv[3])) if len(sys.argv) > 3 else None
EXTRA = {'cfg4sparse': (dict(dynamics_type='double_integrator', num_agents=32, num_landmarks=2, use_safety_filter=True, world_size=40, episode_length=250), {}, 8192, 6249),
         'cfg2sparse': (dict(dynamics_type='double_integrator', num_agents=8, num_landmarks=2, use_safety_filter=True, world_size=20, episode_length=250), {}, 4096, 6249)}
if wl in EXTRA:
    kw, flags, n, episode = EXTRA[wl]; args, flags = G.default_args(**kw), G.BinaryFlags(flags)
else:
    args, flags, n, episode = B.build_args(wl)
env = B200GraphVecEnv(args, num_envs=n, seed=1234, binary_cfg=flags, tuning=tuning)
if mode != 'dense':
    env.enable_edge_output(dense_adj=(mode == 'dense+edges'))
acts = torch.randint(0, 25, (6, n, env.N), device='cuda', dtype=torch.int32)
env.reset(episode)
for t in range(6):
    env.step(acts[t], episode)
torch.cuda.synchronize()
print('ok', env.launch_info())
