#!/usr/bin/env python
"""Golden fixture for SURVEY 8a row a16: the renderer's world graph, `SafeAamScenario.update_graph`
(/root/reference/multiagent/custom_scenarios/navigation_graph_safe.py:996-1015), recorded from the UNMODIFIED reference.

For two rollouts (double integrator and airtaxi, agents steered to their goals so that landmarks and agents get
disconnected) the script records, after every env.step, the simulator state and what `update_graph` - which the next
env.step calls first (environment.py:964-965) - writes into world.edge_list / world.edge_weight; plus one crafted state
with two entities EXACTLY max_edge_dist apart (the radius test is inclusive here, quirk Q7).
-> tests/golden/aux/world_graph.npz        usage: python oracle/gen_world_graph_golden.py   (build container only)
`python oracle/gen_world_graph_golden.py obstacles` records one rollout of the DECLARED obstacle extension instead
(ref_harness._obstacle_extension; update_graph itself is unmodified reference code) -> aux/world_graph_obstacles.npz
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as H          # noqa: E402
import gen_golden as GG          # noqa: E402

KEYS = ('agent_values', 'done', 'reached_goal', 'landmark_pos', 'landmark_heading', 'landmark_speed')
OBSTACLES = len(sys.argv) > 1 and sys.argv[1] == 'obstacles'


def record(env, sc, out, tag, t):
    s = H.snapshot(env)
    sc.update_graph(env.world)                                   # reads world.cached_dist_mag as the next step would
    for k in KEYS + (('obstacle_pos', 'num_obstacle_collisions') if OBSTACLES else ()):
        out.setdefault(f'{tag}__{k}', []).append(np.asarray(s[k]))
    out.setdefault(f'{tag}__edge_list', []).append(np.asarray(env.world.edge_list, dtype=np.int64))
    out.setdefault(f'{tag}__edge_weight', []).append(np.asarray(env.world.edge_weight, dtype=np.float64))


def main():
    out, meta = {}, {}
    cases = (('di4', dict(dynamics_type='double_integrator', num_agents=4, world_size=3, episode_length=60), 45),
             ('at4', dict(dynamics_type='airtaxi', num_agents=4, world_size=6, episode_length=120), 60))
    if OBSTACLES:
        cases = (('di4_obst3', dict(dynamics_type='double_integrator', num_agents=4, world_size=3, episode_length=60, num_obstacles=3), 45),)
    for tag, kw, T in cases:
        args = H.make_args(num_landmarks=2, use_safety_filter=False, **kw)
        env = H.make_env(args, seed=11, obstacle_extension=OBSTACLES)
        sc = H.scenario_of(env)
        env.reset(0)
        rng = np.random.default_rng(5)
        N = args.num_agents
        for i in range(N):
            GG.inject_near_goal(env, sc, i, back=0.15 * (i + 1) if tag == 'di4' else 0.4 * (i + 1))
        env.world.calculate_distances()
        record(env, sc, out, tag, -1)
        for t in range(T):
            a = [GG.greedy_action(env, sc, i, rng) for i in range(N)]
            env.step(H.one_hot(a))
            record(env, sc, out, tag, t)
        meta[tag] = dict(args=kw, steps=T + 1, reached_max=int(np.max(out[f'{tag}__reached_goal'])),
                         done_any=bool(np.any(out[f'{tag}__done'])))
    # crafted: entity 1 exactly max_edge_dist (4.0) from entity 0, entity 2 one ulp farther
    if OBSTACLES:
        return write(out, meta, 'world_graph_obstacles.npz')
    args = H.make_args(num_landmarks=2, use_safety_filter=False, dynamics_type='double_integrator', num_agents=3, world_size=4)
    env = H.make_env(args, seed=3)
    sc = H.scenario_of(env)
    env.reset(0)
    w = env.world
    w.agents[0].state.values = np.array([0.0, 0.0, 0.0, 0.0])
    w.agents[1].state.values = np.array([4.0, 0.0, 0.0, 0.0])
    w.agents[2].state.values = np.array([0.0, np.nextafter(4.0, 5.0), 0.0, 0.0])
    w.calculate_distances()
    record(env, sc, out, 'boundary', 0)
    meta['boundary'] = dict(args=dict(dynamics_type='double_integrator', num_agents=3, world_size=4), steps=1)
    write(out, meta, 'world_graph.npz')


def write(out, meta, fname):
    packed = {'meta': np.array(json.dumps(meta))}
    for k, v in out.items():
        if k.endswith('edge_list') or k.endswith('edge_weight'):
            packed[k + '_len'] = np.array([x.shape[-1] for x in v])
            packed[k] = np.concatenate(v, axis=-1)
        else:
            packed[k] = np.stack(v)
    path = os.path.join(os.path.dirname(HERE), 'tests', 'golden', 'aux', fname)
    np.savez_compressed(path, **packed)
    print('wrote', path, os.path.getsize(path), meta)


if __name__ == '__main__':
    main()
