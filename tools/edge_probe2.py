import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'tests'))
import torch
import bench as B
from layered_safe_marl_b200 import B200GraphVecEnv
wl, mode = sys.argv[1], sys.argv[2]
args, flags, n, episode = B.build_args(wl)
env = B200GraphVecEnv(args, num_envs=n, seed=1234, binary_cfg=flags, tuning=dict(chunks=1))
if mode != 'dense':
    env.enable_edge_output(dense_adj=(mode == 'dense+edges'))
acts = torch.randint(0, 25, (6, n, env.N), device='cuda', dtype=torch.int32)
env.reset(episode)
for t in range(6):
    env.step(acts[t], episode)
torch.cuda.synchronize()
print('ok')
