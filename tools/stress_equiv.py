#!/usr/bin/env python
"""Long back-to-back rollouts (auto-resets included) under every launch arrangement must end in bit-identical states:
a race in the dependency-less late pair kernel or in the chunked fork/join would show up as a divergence.
usage: tools/stress_equiv.py [steps]"""
import hashlib, os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'tests'))
import torch
import _golden as G
from layered_safe_marl_b200 import B200GraphVecEnv

T = int(sys.argv[1]) if len(sys.argv) > 1 else 400
cases = [('di8', dict(num_agents=8, world_size=4), 4096), ('air10', dict(dynamics_type='airtaxi', num_agents=10, world_size=6), 4096),
         ('di32', dict(num_agents=32, world_size=4), 512)]
for name, kw, n in cases:
    args = G.default_args(use_safety_filter=True, episode_length=30, **kw)
    digests = {}
    # pair_placement: 2 front, 0 late, 3 middle (lsm_tuning); chunks: env ranges on library-owned streams
    for placement, chunks in ((2, 1), (0, 1), (3, 1), (0, 4), (3, 4)):
        env = B200GraphVecEnv(args, num_envs=n, seed=17, tuning=dict(pair_placement=placement, chunks=chunks))
        gen = torch.Generator(device=env.device); gen.manual_seed(5)
        env.reset(6249)
        h = hashlib.sha256()
        filt = 0
        for t in range(T):
            out = env.step(torch.randint(0, 25, (n, env.N), generator=gen, device=env.device, dtype=torch.int32), 6249)
            if t % 50 == 49 or t == T - 1:
                for x in out[:6]:
                    h.update(x.cpu().numpy().tobytes())
                st = env.get_state()
                filt += int(st['safety_filtered'].sum())
                for k in sorted(st):
                    h.update(np.ascontiguousarray(st[k]).tobytes())
        digests[(placement, chunks)] = (h.hexdigest()[:16], filt, env.launch_info()['chunks'])
        env.close()
    ok = len({d[0] for d in digests.values()}) == 1
    print(name, 'IDENTICAL' if ok else 'DIVERGED', digests, flush=True)
    assert ok
