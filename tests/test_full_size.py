"""GPU, BASELINE.json's full sizes (4096 x 8, 16384 x 10 airtaxi, 8192 x 32): size-independent properties of the
graph observation, idempotence of the emission kernel, and the C oracle on a random SAMPLE of the environments of the
full-size batch (states copied out of the device batch after a warm-up, same actions, one more step)."""
import os
import sys

import numpy as np
import pytest

import _golden as G

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle'))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

pytestmark = pytest.mark.gpu

FULL = {   # name: (workload key of bench.py, warm-up steps)
    'cfg2': 6, 'cfg3': 4, 'cfg4': 3,
    'cfg3_obst4': 4,       # BASELINE configs[2] with its obstacles (declared extension)
}


def _make(workload):
    import torch
    import bench as B
    from layered_safe_marl_b200 import B200GraphVecEnv
    args, flags, n_envs, episode = B.build_args(workload)
    env = B200GraphVecEnv(args, num_envs=n_envs, seed=77, binary_cfg=flags)
    return torch, env, args, flags, n_envs, episode


@pytest.mark.parametrize('workload', sorted(FULL))
def test_full_size_graph_properties_and_sampled_oracle(workload):
    import oracle_env as O
    from layered_safe_marl_b200 import config as cfg
    torch, env, args, flags, n, episode = _make(workload)
    N, E, F = env.N, env.E, env.F
    gen = torch.Generator(device=env.device); gen.manual_seed(3)
    env.reset(episode)
    for _ in range(FULL[workload]):
        env.step(torch.randint(0, 25, (n, N), generator=gen, device=env.device, dtype=torch.int32), episode)
    # ---- state of a random sample BEFORE the checked step -> oracle
    rng = np.random.default_rng(1)
    sample = np.sort(rng.choice(n, size=256, replace=False))
    s_all = env.get_state()
    s_smp = {k: np.asarray(v)[sample] for k, v in s_all.items()}
    params = cfg.scenario_params_from_args(args, binary_cfg=flags)
    vg, tg = G.value_grid_for(params)
    ora = O.OracleEnv(params.asdict(), len(sample), value_grid=vg, ttr_grid=tg, seed=77, nthreads=8)
    ora.set_state(s_smp)
    acts = torch.randint(0, 25, (n, N), generator=gen, device=env.device, dtype=torch.int32)
    obs, aid, node_obs, adj, rew, done, _ = env.step(acts, episode)
    torch.cuda.synchronize()
    # the specialised airtaxi pipeline evaluates the relative position in its rotation form (see oracle/lsm_oracle.c)
    prev_form = O.set_relative_state_form(1 if (params.dynamics != 0 and env.launch_info()['specialised'] == 1) else 0)
    try:
        ora.step(acts.cpu().numpy()[sample], episode=episode, auto_reset=False)
    finally:
        O.set_relative_state_form(prev_form)

    # ---- properties on the device, whole batch
    r = float(params.coordination_range)
    a = adj
    assert torch.equal(a, a.transpose(-1, -2)), "adjacency of every observer must be symmetric"
    assert bool((torch.diagonal(a, dim1=-2, dim2=-1) == 0).all()), "zero diagonal"
    assert bool(((a == 0) | ((a > 0) & (a < r))).all()), "entries are 0 or a distance strictly inside the radius"
    # type flag: agents 0, landmarks 1, obstacles 2 (last node feature)
    NM = N + env.M
    assert bool((node_obs[:, :, :N, F - 1] == 0).all()) and bool((node_obs[:, :, N:NM, F - 1] == 1).all())
    assert bool((node_obs[:, :, NM:, F - 1] == 2).all())
    # observer i's own node: zero relative position; row i of its adjacency is the norm of the relative positions
    idx = torch.arange(N, device=env.device)
    own = node_obs[:, idx, idx, :2]
    assert bool((own == 0).all())
    rel = node_obs[..., :2].double()
    dist = torch.sqrt((rel * rel).sum(-1))                       # (n, N, E): |p_e - p_i| (rotations keep the norm)
    row = a[:, idx, idx, :].double()                              # (n, N, E): adjacency row of the observer itself
    on = row != 0
    assert bool((torch.abs(row - dist)[on] <= 2e-6 + 1e-5 * dist[on]).all()), "adjacency row vs node-feature positions"
    # reward range and done flags
    assert bool((rew >= params.min_reward - 1e-6).all()) and bool((rew <= params.max_reward + 1e-6).all())

    # ---- the emission kernel is idempotent (same records -> bit-identical tiles)
    n0, a0 = node_obs.clone(), adj.clone()
    env.emit_only()
    torch.cuda.synchronize()
    assert torch.equal(env.node_obs, n0) and torch.equal(env.adj, a0)

    # ---- sampled oracle parity at full size
    smp = torch.as_tensor(sample, device=env.device)
    so, sc_all = ora.get_state(), env.get_state()
    bad = np.zeros(len(sample), dtype=bool)
    for k in ('reached_goal', 'done', 'safety_filtered', 'deconflicting_agent_index', 'num_agent_collisions'):
        bad |= (np.asarray(so[k]) != np.asarray(sc_all[k])[sample]).reshape(len(sample), -1).any(axis=1)
    # environments that auto-reset on the device in this step are not comparable to the (non-resetting) oracle sample
    reset_now = env.env_i32[3].cpu().numpy()[sample].astype(bool)
    bad_cmp = bad & ~reset_now
    assert bad_cmp.sum() == 0, f"{bad_cmp.sum()} sampled envs differ from the oracle in a discrete output"
    ok = ~bad & ~reset_now
    assert ok.sum() >= len(sample) - 16
    okt = torch.as_tensor(np.nonzero(ok)[0], device=env.device)
    G.assert_same_mask((adj[smp][okt] != 0).cpu().numpy(), ora.adj[ok] != 0, 'sampled adj pattern')
    G.assert_close(adj[smp][okt].cpu().numpy(), ora.adj[ok], 'sampled adj')
    G.assert_close(node_obs[smp][okt].cpu().numpy(), ora.node_obs[ok], 'sampled node_obs')
    G.assert_close(obs[smp][okt].cpu().numpy(), ora.obs[ok], 'sampled obs')
    G.assert_close(rew[smp][okt].cpu().numpy(), ora.reward[ok], 'sampled reward')
    assert np.array_equal(np.asarray(sc_all['agent_values'])[sample][ok], np.asarray(so['agent_values'])[ok]), 'sampled float64 states must be bit-identical'
    env.close()
