"""Generate the golden rollouts under tests/golden/ by running the UNMODIFIED reference.

    python oracle/gen_golden.py            # needs /root/reference (build container only)

TEST INFRASTRUCTURE ONLY. Each fixture `tests/golden/<case>.npz` holds, for one environment:
  meta            json: make_world args, RewardBinaryConfig switches, episode index, T
  s0_<key>        full simulator state right before the first step (oracle/ref_harness.snapshot)
  obs0/node_obs0/adj0   what env.reset() would return for that state
  actions         (T, N) discrete action indices fed to env.step (as one-hot, like the runner)
  obs/node_obs/adj/reward/done      per-step outputs of MultiAgentGraphEnv.step (float32 / bool)
  st_<key>        per-step post-step state (T, ...)
  info_<key>      per-step info dict entries (T, N)
  ep_info         the 8-key summary env.reset() reports for the rolled-out episode

The filter-off double-integrator cases are pure reference code. Filter-on, HJ_VALUE and airtaxi cases
run the reference's control flow over oracle/ref_stubs' restated third-party arithmetic and the
synthetic grids of layered_safe_marl_b200.hj_grid ("restated oracle, synthetic grid").
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as H  # noqa: E402

OUT_DIR = os.environ.get('LSM_GOLDEN_OUT') or os.path.join(os.path.dirname(HERE), 'tests', 'golden')   # override: reproducibility test


def greedy_action(env, sc, i, rng):
    """A crude goal-seeking choice among the 25 motion primitives (drives goal/done events)."""
    world = env.world
    agent = world.agents[i]
    goal = sc.get_agent_current_goal(agent, world)
    gp, gh, gs = goal.state.p_pos, goal.heading, goal.speed
    p = agent.state.p_pos
    d = gp - p
    dist = np.linalg.norm(d)
    if env.dynamics_type.name == 'DoubleIntegratorXY':
        if dist < 0.6:
            vdes = gs * np.array([np.cos(gh), np.sin(gh)]) + 0.5 * d
        else:
            vdes = 0.5 * d / max(dist, 1e-9)
        best, bi = None, 0
        opts = np.linspace(-0.5, 0.5, 5)
        for ix in range(5):
            for iy in range(5):
                vn = agent.state.p_vel + 0.1 * np.array([opts[ix], opts[iy]])
                c = np.linalg.norm(vn - vdes)
                if best is None or c < best:
                    best, bi = c, ix * 5 + iy
        return bi
    th, v = agent.state.theta, agent.state.speed
    bearing = np.arctan2(d[1], d[0]) if dist > 0.5 else gh
    err = np.arctan2(np.sin(bearing - th), np.cos(bearing - th))
    w_opts = np.linspace(-0.1, 0.1, 5)
    a_opts = np.linspace(-0.001, 0.002, 5)
    ir = int(np.argmin(np.abs(w_opts - np.clip(err, -0.1, 0.1))))
    ia = int(np.argmin(np.abs(a_opts - np.clip(gs - v, -0.001, 0.002))))
    return ir * 5 + ia


def inject_near_goal(env, sc, i, back, reached=0, lateral=0.0):
    """Put agent i `back` behind its (reached-th) goal, aligned with the goal heading/speed."""
    world = env.world
    sc.reached_goal[i] = reached
    agent = world.agents[i]
    goal = sc.get_agent_current_goal(agent, world)
    gh, gs = goal.heading, goal.speed
    dirv = np.array([np.cos(gh), np.sin(gh)])
    lat = np.array([-np.sin(gh), np.cos(gh)])
    pos = goal.state.p_pos - back * dirv + lateral * lat
    if env.dynamics_type.name == 'DoubleIntegratorXY':
        agent.state.values = np.array([pos[0], pos[1], gs * dirv[0], gs * dirv[1]])
    else:
        agent.state.values = np.array([pos[0], pos[1], gh, gs])


def inject_pair_conflict(env, i, j, gap, closing):
    """Put agent j `gap` away from agent i, closing head-on (exercises bang-bang / QP branches)."""
    world = env.world
    a, b = world.agents[i], world.agents[j]
    if env.dynamics_type.name == 'DoubleIntegratorXY':
        pa = a.state.p_pos.copy()
        a.state.values = np.array([pa[0], pa[1], closing, 0.0])
        b.state.values = np.array([pa[0] + gap, pa[1] + 0.03, -closing, 0.0])
    else:
        pa = a.state.p_pos.copy()
        a.state.values = np.array([pa[0], pa[1], 0.0, closing])
        b.state.values = np.array([pa[0] + gap, pa[1] + 0.05, np.pi, closing])


def observe_now(env):
    env.world.calculate_distances()
    obs, nobs, adj = [], [], []
    for agent in env.agents:
        obs.append(env._get_obs(agent))
        n, a = env._get_graph_obs(agent)
        nobs.append(n)
        adj.append(a)
    env.world.calculate_distances()   # undo the in-place masking for the rollout that follows
    return np.array(obs), np.array(nobs), np.array(adj)


def run_case(name, args_kw, flags, episode, T, seed, greedy_agents=(), injections=(), post=None):
    interp_float32 = bool(args_kw.get('interp_float32', False))     # not a reference argument: selects the stub's arithmetic
    obstacle_extension = bool(args_kw.get('obstacle_extension', False))   # not a reference argument: ref_harness._obstacle_extension
    args = H.make_args(**{k: v for k, v in args_kw.items() if k not in ('interp_float32', 'obstacle_extension')})
    env = H.make_env(args, seed=seed, interp_float32=interp_float32, obstacle_extension=obstacle_extension, **flags)
    sc = H.scenario_of(env)
    N = args.num_agents
    env.reset(episode)
    for inj in injections:
        kind = inj[0]
        if kind == 'near_goal':
            inject_near_goal(env, sc, *inj[1:])
        elif kind == 'pair':
            inject_pair_conflict(env, *inj[1:])
        elif kind == 'obstacle_at_agent':      # obstacle k just ahead of agent i (drives Num_obst_collisions)
            k, i, ahead = inj[1:]
            a = env.world.agents[i]
            env.world.obstacles[k].state.p_pos = a.state.p_pos + ahead * np.array([np.cos(a.state.theta), np.sin(a.state.theta)])
    for agent in env.world.agents:   # keep min_time consistent with injected positions (reset bookkeeping)
        sc.min_time(agent, env.world)
    s0 = H.snapshot(env)
    obs0, nobs0, adj0 = observe_now(env)
    rng = np.random.default_rng(seed + 1000)
    rec = {k: [] for k in ('actions', 'obs', 'node_obs', 'adj', 'reward', 'done')}
    st = {}
    info = {}
    for t in range(T):
        a = rng.integers(0, 25, N)
        for i in greedy_agents:
            if rng.random() < 0.85:
                a[i] = greedy_action(env, sc, i, rng)
        obs, aid, nobs, adj, rew, done, infos = env.step(H.one_hot(a))
        rec['actions'].append(a.astype(np.int32))
        rec['obs'].append(np.array(obs, dtype=np.float32))
        rec['node_obs'].append(np.array(nobs, dtype=np.float32))
        rec['adj'].append(np.array(adj, dtype=np.float32))
        rec['reward'].append(np.array(rew, dtype=np.float64).reshape(N))
        rec['done'].append(np.array(done, dtype=bool))
        snap = H.snapshot(env)
        for k, v in snap.items():
            st.setdefault(k, []).append(np.array(v))
        ia = H.info_arrays(infos, N)
        for k, v in ia.items():
            info.setdefault(k, []).append(v)
        assert np.array_equal(np.array(aid).reshape(-1), np.arange(N))
    # the summary the next reset reports for this episode (environment.py:1046-1074)
    _, _, _, _, ep_info = env.reset(episode)
    ep = np.array([ep_info[k] for k in ('travel_time_mean', 'travel_distance_mean', 'done_percentage',
                                        'num_reached_goal_mean', 'conflict_percentage', 'min_distance_mean',
                                        'min_distance_min', 'multiple_engagement_percentage')], dtype=np.float64)
    out = {'meta': np.array(json.dumps(dict(name=name, args=args_kw, flags=flags, episode=int(episode), T=int(T),
                                            seed=int(seed), num_total_episode=int(sc.num_total_episode))))}
    for k, v in s0.items():
        out['s0_' + k] = np.array(v)
    out['obs0'] = obs0.astype(np.float32)
    out['node_obs0'] = nobs0.astype(np.float32)
    out['adj0'] = adj0.astype(np.float32)
    for k, v in rec.items():
        out[k] = np.stack(v)
    for k, v in st.items():
        out['st_' + k] = np.stack(v)
    for k, v in info.items():
        out['info_' + k] = np.stack(v)
    out['ep_info'] = ep
    os.makedirs(OUT_DIR, exist_ok=True)
    path = os.path.join(OUT_DIR, name + '.npz')
    np.savez_compressed(path, **out)
    n_done = int(out['st_done'][-1].sum())
    n_filt = int(out['st_safety_filtered'].sum())
    n_reach = int(out['st_reached_goal'][-1].sum())
    print(f"{name:28s} T={T:3d} N={N:2d} reached={n_reach} agent_done={n_done} filtered_events={n_filt} "
          f"rew[min,max]=[{out['reward'].min():.2f},{out['reward'].max():.2f}] "
          f"size={os.path.getsize(path) / 1024:.0f} KiB")


DI = dict(dynamics_type='double_integrator')
AT = dict(dynamics_type='airtaxi', world_size=6)
ALL_FLAGS = dict(SAFETY_VIOLATION=True, HJ_VALUE=True, POTENTIAL_CONFLICT=True,
                 SEPARATION_DISTANCE_CURRICULUM=True, INITIAL_PHASE_USE_SAFETY_FILTER=True,
                 DIFF_FROM_FILTERED_ACTION=True)

CASES = [
    # BASELINE config 1 analogue: pure reference code, magnetic-field reward active (episode 0)
    dict(name='di3_nofilter_ep0', args_kw=dict(DI, num_agents=3, episode_length=25), flags={}, episode=0, T=30, seed=1),
    dict(name='di3_nofilter_goals', args_kw=dict(DI, num_agents=3, episode_length=50), flags={}, episode=3000, T=45,
         seed=2, greedy_agents=(0, 1, 2),
         injections=(('near_goal', 0, 0.45, 0), ('near_goal', 1, 0.40, 1, 0.05), ('near_goal', 2, 0.9, 0, -0.1))),
    dict(name='di8_filter', args_kw=dict(DI, num_agents=8, use_safety_filter=True, world_size=2, episode_length=250),
         flags={}, episode=6249, T=30, seed=3, greedy_agents=(0, 1, 2, 3),
         injections=(('near_goal', 0, 0.5, 1), ('pair', 4, 5, 0.9, 0.45), ('pair', 6, 7, 1.6, 0.3))),
    dict(name='di8_filter_allflags', args_kw=dict(DI, num_agents=8, use_safety_filter=True, world_size=2,
                                                  episode_length=250), flags=ALL_FLAGS, episode=3500, T=30, seed=4,
         greedy_agents=(0, 1, 2), injections=(('near_goal', 1, 0.45, 1), ('pair', 3, 4, 0.7, 0.4), ('pair', 5, 6, 1.2, 0.45))),
    dict(name='di4_collab_conflict', args_kw=dict(DI, num_agents=4, collaborative=True, world_size=1, episode_length=25),
         flags=dict(SAFETY_VIOLATION=True, POTENTIAL_CONFLICT=True), episode=5000, T=28, seed=5, greedy_agents=(0,),
         injections=(('near_goal', 0, 0.4, 1), ('pair', 1, 2, 0.5, 0.3))),
    dict(name='di8_filter_ep0', args_kw=dict(DI, num_agents=8, use_safety_filter=True, episode_length=250), flags={},
         episode=0, T=8, seed=6),
    dict(name='di3_internal2', args_kw=dict(DI, num_agents=3, use_safety_filter=True, num_internal_step=2, world_size=1,
                                            episode_length=250), flags=dict(DIFF_FROM_FILTERED_ACTION=True),
         episode=6249, T=20, seed=7, injections=(('pair', 0, 1, 0.8, 0.4),)),
    dict(name='di16_dense', args_kw=dict(DI, num_agents=16, use_safety_filter=True, world_size=4, episode_length=250),
         flags=dict(POTENTIAL_CONFLICT=True), episode=6249, T=6, seed=8, greedy_agents=(0, 1),
         injections=(('near_goal', 0, 0.4, 1),)),
    dict(name='at4_nofilter_ep0', args_kw=dict(AT, num_agents=4, episode_length=30), flags={}, episode=0, T=34, seed=9,
         greedy_agents=(0,), injections=(('near_goal', 0, 0.5, 1),)),
    # BASELINE config 3 analogue (obstacle-free, as the reference itself cannot run obstacles)
    dict(name='at10_filter_pc', args_kw=dict(AT, num_agents=10, use_safety_filter=True, episode_length=350),
         flags=dict(POTENTIAL_CONFLICT=True), episode=6249, T=36, seed=10, greedy_agents=(0, 1, 2, 3, 4),
         injections=(('near_goal', 0, 0.6, 1), ('near_goal', 1, 0.7, 0, 0.05), ('pair', 5, 6, 2.0, 0.08),
                     ('pair', 7, 8, 3.5, 0.06))),
    # graph_feat_type='global' (navigation_graph_safe.py:1017-1036): observer-independent 7-wide node features
    dict(name='di3_global_feat', args_kw=dict(DI, num_agents=3, episode_length=50, graph_feat_type='global'), flags={},
         episode=3000, T=30, seed=12, greedy_agents=(0, 1), injections=(('near_goal', 0, 0.45, 1), ('near_goal', 1, 0.40, 0, 0.05))),
    dict(name='at4_global_feat', args_kw=dict(AT, num_agents=4, episode_length=30, graph_feat_type='global'), flags={},
         episode=0, T=12, seed=13, greedy_agents=(0,), injections=(('near_goal', 0, 0.5, 1),)),
    # the second declared interpolation arithmetic (float32, jax default dtype): same scenarios as di8_filter / at10_filter_pc
    dict(name='di8_filter_f32', args_kw=dict(DI, num_agents=8, use_safety_filter=True, world_size=2, episode_length=250,
                                             interp_float32=True),
         flags={}, episode=6249, T=30, seed=3, greedy_agents=(0, 1, 2, 3),
         injections=(('near_goal', 0, 0.5, 1), ('pair', 4, 5, 0.9, 0.45), ('pair', 6, 7, 1.6, 0.3))),
    dict(name='at10_filter_pc_f32', args_kw=dict(AT, num_agents=10, use_safety_filter=True, episode_length=350, interp_float32=True),
         flags=dict(POTENTIAL_CONFLICT=True), episode=6249, T=36, seed=10, greedy_agents=(0, 1, 2, 3, 4),
         injections=(('near_goal', 0, 0.6, 1), ('near_goal', 1, 0.7, 0, 0.05), ('pair', 5, 6, 2.0, 0.08),
                     ('pair', 7, 8, 3.5, 0.06))),
    # DECLARED obstacle extension (the reference raises for num_obstacles > 0; ref_harness._obstacle_extension completes the
    # two raising statements, everything else is the reference's own obstacle code): 'relative' and 'global' features, both
    # dynamics, BASELINE config 3's shape '+ obstacles' (airtaxi, 10 agents, POTENTIAL_CONFLICT, filter on, 4 obstacles)
    dict(name='di3_obst2', args_kw=dict(DI, num_agents=3, num_obstacles=2, episode_length=50, obstacle_extension=True), flags={},
         episode=3000, T=30, seed=21, greedy_agents=(0, 1), injections=(('near_goal', 0, 0.45, 1), ('obstacle_at_agent', 0, 2, 0.09),
                                                                       ('obstacle_at_agent', 1, 0, 0.25))),
    dict(name='di4_obst3_filter_global', args_kw=dict(DI, num_agents=4, num_obstacles=3, use_safety_filter=True, world_size=2,
                                                       episode_length=250, graph_feat_type='global', obstacle_extension=True),
         flags={}, episode=6249, T=20, seed=22, greedy_agents=(0,), injections=(('pair', 1, 2, 0.9, 0.4), ('obstacle_at_agent', 0, 3, 0.15))),
    dict(name='at10_obst4_filter_pc', args_kw=dict(AT, num_agents=10, num_obstacles=4, use_safety_filter=True, episode_length=350,
                                                   obstacle_extension=True),
         flags=dict(POTENTIAL_CONFLICT=True), episode=6249, T=24, seed=23, greedy_agents=(0, 1, 2),
         injections=(('near_goal', 0, 0.6, 1), ('pair', 5, 6, 2.0, 0.08), ('obstacle_at_agent', 0, 3, 0.2), ('obstacle_at_agent', 3, 7, 0.15))),
    # edge shapes: a single agent (no "others": the filter, min distance and engagement terms see empty lists), and more than two
    # landmarks per agent (goal chains of 3 / 4, run-time L on the generic kernel)
    dict(name='di1_single_filter', args_kw=dict(DI, num_agents=1, use_safety_filter=True, episode_length=60), flags=ALL_FLAGS,
         episode=6249, T=40, seed=31, greedy_agents=(0,), injections=(('near_goal', 0, 0.5, 0),)),
    dict(name='di2_landmarks4', args_kw=dict(DI, num_agents=2, num_landmarks=4, use_safety_filter=True, episode_length=120, world_size=2),
         flags={}, episode=6249, T=70, seed=32, greedy_agents=(0, 1), injections=(('near_goal', 0, 0.4, 0), ('near_goal', 1, 0.5, 2))),
    dict(name='at3_landmarks3', args_kw=dict(AT, num_agents=3, num_landmarks=3, use_safety_filter=True, episode_length=200),
         flags=dict(POTENTIAL_CONFLICT=True), episode=6249, T=40, seed=33, greedy_agents=(0, 1, 2),
         injections=(('near_goal', 0, 0.5, 0), ('near_goal', 1, 0.6, 1), ('near_goal', 2, 0.4, 2))),
    dict(name='at6_allflags', args_kw=dict(AT, num_agents=6, use_safety_filter=True, world_size=3, episode_length=350),
         flags=ALL_FLAGS, episode=4000, T=30, seed=11, greedy_agents=(0, 1),
         injections=(('near_goal', 0, 0.5, 1), ('pair', 2, 3, 1.5, 0.085), ('pair', 4, 5, 2.5, 0.035))),
]


def main():
    only = set(sys.argv[1:])
    for c in CASES:
        if only and c['name'] not in only:
            continue
        run_case(**c)


if __name__ == '__main__':
    main()
