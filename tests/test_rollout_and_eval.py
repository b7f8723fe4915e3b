"""N1 / N4 rows: device rollout storage, GraphDummyVecEnv surface, deterministic eval scenarios.

CPU tests check the host logic against numpy restatements of the reference lines they mirror and run the
eval scenarios through the C oracle; GPU tests check the CUDA path (zero-copy slot binding, parity with the oracle)."""
import os
import sys
import types

import numpy as np
import pytest
import torch

import _golden as G

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle'))

from layered_safe_marl_b200 import config as cfg, eval_scenarios as ES
from layered_safe_marl_b200.rollout import DeviceGraphRolloutBuffer
from layered_safe_marl_b200.spaces import Discrete


def _fake_env(n, N, L, D, F, device='cpu'):
    e = types.SimpleNamespace()
    e.num_envs, e.N, e.E, e.D, e.F, e.device = n, N, N * (1 + L), D, F, torch.device(device)
    e.action_space = [Discrete(25) for _ in range(N)]
    e.agent_id = torch.arange(N, dtype=torch.int32).view(1, N, 1).expand(n, N, 1).contiguous()
    return e


def _ref_insert(buf, step, obs, agent_id, node_obs, adj, rewards, dones, n, N):
    """numpy restatement of GMPERunner.insert (graph_mpe_runner.py:444-487) + GraphReplayBuffer.insert
    (graph_buffer.py:223-251) for the fields the environment produces."""
    dones_env = np.all(dones, axis=1)
    masks = np.ones((n, N, 1), dtype=np.float32)
    masks[dones] = 0.0
    active = np.ones((n, N, 1), dtype=np.float32)
    active[dones] = 0.0
    active[dones_env] = 1.0
    share_obs = np.expand_dims(obs.reshape(n, -1), 1).repeat(N, axis=1)
    share_agent_id = np.expand_dims(agent_id.reshape(n, -1), 1).repeat(N, axis=1)
    buf['share_obs'][step + 1] = share_obs; buf['obs'][step + 1] = obs; buf['node_obs'][step + 1] = node_obs
    buf['adj'][step + 1] = adj; buf['agent_id'][step + 1] = agent_id; buf['share_agent_id'][step + 1] = share_agent_id
    buf['rewards'][step] = rewards; buf['masks'][step + 1] = masks; buf['active_masks'][step + 1] = active


def test_rollout_buffer_matches_reference_insert_and_gae():
    n, N, L, D, F, T = 5, 3, 2, 7, 10, 6
    E = N * (1 + L)
    env = _fake_env(n, N, L, D, F)
    rb = DeviceGraphRolloutBuffer(env, episode_length=T, gamma=0.97, gae_lambda=0.9)
    rng = np.random.default_rng(0)
    ref = {k: np.zeros(tuple(getattr(rb, k).shape), dtype=np.float32 if getattr(rb, k).dtype == torch.float32 else np.int32)
           for k in ('share_obs', 'obs', 'node_obs', 'adj', 'agent_id', 'share_agent_id', 'rewards')}
    ref['masks'] = np.ones(tuple(rb.masks.shape), dtype=np.float32)
    ref['active_masks'] = np.ones(tuple(rb.masks.shape), dtype=np.float32)
    agent_id = np.tile(np.arange(N, dtype=np.int32).reshape(1, N, 1), (n, 1, 1))
    # agent ids are constant: the device buffer fills every slot once at construction, the reference fills slot 0 in warmup
    ref['agent_id'][0] = agent_id
    ref['share_agent_id'][0] = np.expand_dims(agent_id.reshape(n, -1), 1).repeat(N, axis=1)
    values = rng.normal(size=(T + 1, n, N, 1)).astype(np.float32)
    for t in range(T):
        obs = rng.normal(size=(n, N, D)).astype(np.float32)
        node_obs = rng.normal(size=(n, N, E, F)).astype(np.float32)
        adj = rng.random(size=(n, N, E, E)).astype(np.float32)
        rewards = rng.normal(size=(n, N, 1)).astype(np.float32)
        dones = rng.random((n, N)) < 0.3
        if t == 2:
            dones[1] = True            # a whole env done: active_masks back to one
        _ref_insert(ref, t, obs, agent_id, node_obs, adj, rewards, dones, n, N)
        rb.insert((torch.from_numpy(obs), torch.from_numpy(agent_id), torch.from_numpy(node_obs), torch.from_numpy(adj),
                   torch.from_numpy(rewards[..., 0]), torch.from_numpy(dones), None), values=torch.from_numpy(values[t]))
    for k, v in ref.items():
        np.testing.assert_array_equal(getattr(rb, k).numpy(), v, err_msg=k)
    # GAE, graph_buffer.py:340-360 (no value normaliser)
    vp = values.copy(); ret = np.zeros_like(vp); gae = 0
    for step in reversed(range(T)):
        delta = ref['rewards'][step] + 0.97 * vp[step + 1] * ref['masks'][step + 1] - vp[step]
        gae = delta + 0.97 * 0.9 * ref['masks'][step + 1] * gae
        ret[step] = gae + vp[step]
    rb.compute_returns(torch.from_numpy(values[T]))
    np.testing.assert_allclose(rb.returns.numpy()[:T], ret[:T], rtol=1e-6, atol=1e-6)
    assert rb.step == 0                      # wrapped around after T inserts
    rb.after_update()
    np.testing.assert_array_equal(rb.obs[0].numpy(), rb.obs[-1].numpy())
    oh = DeviceGraphRolloutBuffer.one_hot_actions(torch.from_numpy(rng.integers(0, 25, (n, N, 1))))
    assert oh.shape == (n, N, 25) and float(oh.sum()) == n * N


@pytest.mark.parametrize('name,N', [('circular', 6), ('two_vehicle_conflict', 2), ('three_vehicle_conflict', 3)])
def test_eval_scenarios_run_through_the_oracle(name, N):
    import oracle_env as O
    dyn = 'double_integrator' if name == 'circular' else 'airtaxi'
    kw = dict(num_agents=N, dynamics_type=dyn) if name == 'circular' else {}
    s = ES.build(name, **kw)
    assert s['agent_values'].shape == (1, N, 4) and s['landmark_pos'].shape == (1, 2 * N, 2)
    mind = {}
    for use_filter in (False, True):
        args = G.default_args(dynamics_type=dyn, num_agents=N, use_safety_filter=use_filter, episode_length=400,
                              world_size=4 if dyn == 'double_integrator' else 6)
        params = cfg.scenario_params_from_args(args, binary_cfg=G.BinaryFlags({}))
        vg, tg = G.value_grid_for(params)
        ora = O.OracleEnv(params.asdict(), 1, value_grid=vg, ttr_grid=tg, seed=0, nthreads=1)
        ora.set_state({k: v for k, v in s.items() if k != 'note'})
        ora.reset(episode=params.num_total_episode - 1, sample=False)
        m = np.inf
        act = np.full((1, N), 12, dtype=np.int32)          # zero acceleration / zero turn-rate primitive
        for t in range(60 if dyn == 'airtaxi' else 80):
            ora.step(act, episode=params.num_total_episode - 1, auto_reset=False)
            st = ora.get_state()
            mr = np.asarray(st['min_relative_distance'])
            m = min(m, float(mr[np.isfinite(mr)].min())) if np.isfinite(mr).any() else m
        mind[use_filter] = m
    # the conflict examples are head-on without the filter; the HJ filter (synthetic grid) must not make them closer
    if name != 'circular':
        assert mind[True] >= mind[False] - 1e-9, mind


@pytest.mark.gpu
def test_dummy_vec_env_has_no_auto_reset_and_returns_reset_count():
    from layered_safe_marl_b200 import B200GraphDummyVecEnv
    args = G.default_args(num_agents=3, episode_length=4)
    env = B200GraphDummyVecEnv(args, num_envs=7, seed=3)
    env.reset(0)
    a = torch.zeros((7, 3), dtype=torch.int32, device=env.device)
    for t in range(6):
        out = env.step(a, 0)
        assert len(out) == 8 and out[7] == 0
    st = env.get_state()
    assert (st['current_step'] == 6).all()             # past episode_length: no reset happened
    assert out[5].all()                                # every agent reports done (time limit)


@pytest.mark.gpu
def test_zero_copy_rollout_equals_copying_rollout():
    from layered_safe_marl_b200 import B200GraphVecEnv
    args = G.default_args(num_agents=8, use_safety_filter=True, episode_length=250, world_size=4)
    T = 5
    envs = [B200GraphVecEnv(args, num_envs=64, seed=9) for _ in range(2)]
    bufs = [DeviceGraphRolloutBuffer(envs[0], T, zero_copy=False), DeviceGraphRolloutBuffer(envs[1], T, zero_copy=True)]
    for b in bufs:
        b.warmup(num_current_episode=6249)
    gen = torch.Generator(device='cpu'); gen.manual_seed(1)
    for t in range(T):
        a = torch.randint(0, 25, (64, 8), generator=gen, dtype=torch.int32).to(envs[0].device)
        for env, b in zip(envs, bufs):
            b.insert(env.step(a, 6249), actions=a.unsqueeze(-1).float())
    torch.cuda.synchronize()
    for k in ('obs', 'share_obs', 'node_obs', 'adj', 'rewards', 'masks', 'active_masks', 'actions'):
        assert torch.equal(getattr(bufs[0], k), getattr(bufs[1], k)), k
    assert float(bufs[1].adj.abs().sum()) > 0 and float(bufs[1].node_obs.abs().sum()) > 0


@pytest.mark.gpu
@pytest.mark.parametrize('name,N', [('circular', 8), ('three_vehicle_conflict', 3)])
def test_eval_scenarios_cuda_matches_oracle(name, N):
    import oracle_env as O
    from layered_safe_marl_b200 import B200GraphDummyVecEnv
    dyn = 'double_integrator' if name == 'circular' else 'airtaxi'
    kw = dict(num_agents=N, dynamics_type=dyn) if name == 'circular' else {}
    s = {k: v for k, v in ES.build(name, **kw).items() if k != 'note'}
    args = G.default_args(dynamics_type=dyn, num_agents=N, use_safety_filter=True, episode_length=400,
                          world_size=4 if dyn == 'double_integrator' else 6)
    params = cfg.scenario_params_from_args(args, binary_cfg=G.BinaryFlags({}))
    vg, tg = G.value_grid_for(params)
    ep = params.num_total_episode - 1
    ora = O.OracleEnv(params.asdict(), 1, value_grid=vg, ttr_grid=tg, seed=0, nthreads=1)
    env = B200GraphDummyVecEnv(args, num_envs=1, seed=0)
    ora.set_state(s); ora.reset(episode=ep, sample=False)
    env.set_state(s); env.reset_from_state(ep)
    rng = np.random.default_rng(4)
    for t in range(40):
        a = rng.integers(0, 25, (1, N)).astype(np.int32)
        ora.step(a, episode=ep, auto_reset=False)
        out = env.step(torch.as_tensor(a, device=env.device), ep)
        so, sc = ora.get_state(), env.get_state()
        for k in ('reached_goal', 'done', 'safety_filtered', 'deconflicting_agent_index'):
            np.testing.assert_array_equal(np.asarray(so[k]), np.asarray(sc[k]), err_msg=f't={t} {k}')
        np.testing.assert_array_equal(ora.adj != 0, out[3].cpu().numpy() != 0)
        G.assert_close(out[2].cpu().numpy(), ora.node_obs, f't={t} node_obs')
        G.assert_close(out[4].cpu().numpy(), ora.reward, f't={t} reward')


def _process_adj(adj):
    """TransformerConvNet.process_adj, reference onpolicy/algorithms/utils/gnn.py:376-407 (3-D case), verbatim semantics."""
    batch_size, num_nodes, _ = adj.shape
    edge_index = adj.nonzero(as_tuple=False)
    edge_attr = adj[edge_index[:, 0], edge_index[:, 1], edge_index[:, 2]]
    batch = edge_index[:, 0] * num_nodes
    edge_index = torch.stack([batch + edge_index[:, 1], batch + edge_index[:, 2]], dim=0)
    return edge_index, edge_attr.unsqueeze(1)


@pytest.mark.gpu
@pytest.mark.parametrize('N,dyn', [(8, 'double_integrator'), (3, 'double_integrator'), (10, 'airtaxi')])
def test_edge_list_matches_process_adj(N, dyn):
    """N2: the device COO builder against the reference's own torch code on the adjacency of real steps
    (bit-exact integer indices, bit-exact float32 attributes, ragged graphs incl. empty ones)."""
    from layered_safe_marl_b200 import B200GraphVecEnv
    args = G.default_args(dynamics_type=dyn, num_agents=N, use_safety_filter=False, episode_length=12,
                          world_size=8 if dyn == 'double_integrator' else 12)     # sparse: many pairs beyond the radius
    env = B200GraphVecEnv(args, num_envs=97, seed=5)
    env.reset(0)
    gen = torch.Generator(device='cpu'); gen.manual_seed(2)
    for t in range(14):                                   # crosses an auto-reset
        a = torch.randint(0, 25, (97, N), generator=gen, dtype=torch.int32).to(env.device)
        out = env.step(a, 0)
        adj = out[3]
        ei, ea = env.edge_list()
        ri, ra = _process_adj(adj.reshape(-1, env.E, env.E))
        assert ei.dtype == torch.int64 and ea.dtype == torch.float32 and ea.shape == ra.shape
        assert torch.equal(ei, ri), f't={t}'
        assert torch.equal(ea, ra), f't={t}'
    # an all-zero adjacency (no edges at all) and a user-supplied tensor
    z = torch.zeros_like(env.adj)
    ei, ea = env.edge_list(z)
    assert ei.shape == (2, 0) and ea.shape == (0, 1)
