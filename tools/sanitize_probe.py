#!/usr/bin/env python
"""Small rollouts of every specialised shape for compute-sanitizer (memcheck / racecheck / initcheck).
usage: compute-sanitizer --tool memcheck python tools/sanitize_probe.py"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'tests'))
import torch
import _golden as G
from layered_safe_marl_b200 import B200GraphVecEnv

shapes = [dict(num_agents=8, world_size=4), dict(num_agents=3, world_size=4), dict(num_agents=32, world_size=4),
          dict(dynamics_type='airtaxi', num_agents=10, world_size=6), dict(num_agents=5, world_size=4),
          # declared obstacle extension: the two specialised shapes and two generic ones (maximum obstacle count included)
          dict(dynamics_type='airtaxi', num_agents=10, world_size=6, num_obstacles=4, obstacle_extension=True),
          dict(num_agents=8, world_size=1, num_obstacles=4, obstacle_extension=True),
          dict(num_agents=3, world_size=1, num_obstacles=32, obstacle_extension=True),
          dict(dynamics_type='airtaxi', num_agents=6, world_size=2, num_obstacles=5, obstacle_extension=True)]
if len(sys.argv) > 1 and sys.argv[1] == 'obstacles':
    shapes = shapes[5:]
for kw in shapes:
    for filt in (True, False):
        args = G.default_args(use_safety_filter=filt, episode_length=5, **kw)
        n = 70 if kw['num_agents'] < 32 else 37
        env = B200GraphVecEnv(args, num_envs=n, seed=3)
        env.reset(6249)
        for t in range(8):
            env.step(torch.randint(0, 25, (n, env.N), device=env.device, dtype=torch.int32), 6249)
        env.edge_list()
        env.world_graph()
        env.episode_stats()
        if env.launch_info()['specialised'] == 1:
            env.enable_edge_output(dense_adj=False)
            env.step(torch.randint(0, 25, (n, env.N), device=env.device, dtype=torch.int32), 6249)
            env.edges()
        torch.cuda.synchronize()
        print('ok', kw, 'filter', filt, env.launch_info()['chunks'], flush=True)
        env.close()
