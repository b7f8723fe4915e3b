#!/usr/bin/env python
"""Reference samples for the distributional (Kolmogorov-Smirnov) tests of the reset sampler - SURVEY 8f N3 / row a20.

Draws R resets per configuration from the UNMODIFIED reference (`env.reset(num_current_episode)` ->
`Scenario.random_scenario`, /root/reference/multiagent/custom_scenarios/navigation_graph_safe.py:1199-1367,
`randomly_generate_separated_positions` utils.py:39-68) through oracle/ref_harness.py, reduces every reset to the scalar
features below and stores 1 024 quantiles of each pooled feature (float32) + a few frequencies in
tests/golden/aux/reset_samples.npz. The device sampler (Philox, same draw order) cannot reproduce the reference's
MT19937 stream, so parity is distributional: tests/test_reset_distribution.py compares the same features drawn from
the sampler under test.

usage: python oracle/gen_reset_samples.py             (build container only: needs /root/reference)
       python oracle/gen_reset_samples.py obstacles   -> tests/golden/aux/reset_samples_obstacles.npz: the DECLARED obstacle
           extension (ref_harness._obstacle_extension; placement and the agent-position rejection loop are the reference's
           own code, navigation_graph_safe.py:1204-1249), small worlds with many obstacles so that rejections are frequent
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as H  # noqa: E402

R = 2500
NQ = 1024
CONFIGS = {
    # name: (make_args overrides, curriculum ratio)
    'di_nofilter_r0': (dict(dynamics_type='double_integrator', use_safety_filter=False, world_size=4), 0.0),
    'di_nofilter_r50': (dict(dynamics_type='double_integrator', use_safety_filter=False, world_size=4), 0.5),
    'di_nofilter_r100': (dict(dynamics_type='double_integrator', use_safety_filter=False, world_size=4), 1.0),
    'di_filter_r50': (dict(dynamics_type='double_integrator', use_safety_filter=True, world_size=4), 0.5),
    'at_nofilter_r0': (dict(dynamics_type='airtaxi', use_safety_filter=False, world_size=6), 0.0),
    'at_nofilter_r50': (dict(dynamics_type='airtaxi', use_safety_filter=False, world_size=6), 0.5),
    'at_filter_r100': (dict(dynamics_type='airtaxi', use_safety_filter=True, world_size=6), 1.0),
}
OBSTACLE_CONFIGS = {
    'di_obst12_w1': (dict(dynamics_type='double_integrator', use_safety_filter=False, world_size=1, num_obstacles=12, obstacle_extension=True), 0.5),
    'at_obst16_w1': (dict(dynamics_type='airtaxi', use_safety_filter=True, world_size=1, num_obstacles=16, obstacle_extension=True), 1.0),
}
N, L = 4, 2


def features(agent_values, lpos, lhead, lspeed, dyn, obstacle_pos=None):
    """agent_values (R, N, 4), lpos (R, L*N, 2) [landmark m = order*N + agent], lhead / lspeed (R, L*N) -> dict of pooled
    scalar samples + frequencies. Shared with tests/test_reset_distribution.py (keep in sync: the test imports it)."""
    Rn = agent_values.shape[0]
    g0, g1 = lpos[:, :N], lpos[:, N:2 * N]                       # first / second goal of every agent
    d = g1 - g0
    direction = np.arctan2(d[..., 1], d[..., 0])
    pert = lhead[:, :N] - direction
    pert = np.arctan2(np.sin(pert), np.cos(pert))
    f = {
        'agent_x': agent_values[..., 0].ravel(), 'agent_y': agent_values[..., 1].ravel(),
        'goal0_x': g0[..., 0].ravel(), 'goal0_y': g0[..., 1].ravel(), 'goal1_x': g1[..., 0].ravel(), 'goal1_y': g1[..., 1].ravel(),
        'goal_spacing': np.linalg.norm(d, axis=-1).ravel(),
        'heading_perturbation': pert.ravel(),
        'goal0_speed': lspeed[:, :N].ravel(), 'goal1_speed': lspeed[:, N:2 * N].ravel(),
    }
    if dyn == 'airtaxi':
        f['agent_theta'] = agent_values[..., 2].ravel(); f['agent_speed'] = agent_values[..., 3].ravel()
    if obstacle_pos is not None:      # (R, O, 2)
        f['obstacle_x'] = obstacle_pos[..., 0].ravel(); f['obstacle_y'] = obstacle_pos[..., 1].ravel()
        dao = np.linalg.norm(agent_values[:, :, None, :2] - obstacle_pos[:, None, :, :], axis=-1)
        f['agent_nearest_obstacle'] = dao.min(axis=-1).ravel()      # >= 1.05 * (0.05 + 0.05) by the rejection loop
    freq = {
        # a goal copied from the previous agent (overlap_probability 0.5 per goal, agents 1..N-1)
        'copy_goal0': float(np.mean(np.all(g0[:, 1:] == g0[:, :-1], axis=-1))),
        'copy_goal1': float(np.mean(np.all(g1[:, 1:] == g1[:, :-1], axis=-1))),
        # the last goal keeps the unperturbed direction of the previous leg
        'last_heading_is_direction': float(np.mean(np.abs(np.arctan2(np.sin(lhead[:, N:2 * N] - direction),
                                                                   np.cos(lhead[:, N:2 * N] - direction))) < 1e-9)),
        # double integrator: P(random goal speeds) = min(sloped ratio, 0.8); "fixed" = (max, min)
        'fixed_speed_pattern': float(np.mean((lspeed[:, :N] == lspeed[:, :N].max()) & (lspeed[:, N:2 * N] == lspeed[:, N:2 * N].min()))),
    }
    return f, freq


def main():
    out = {}
    meta = {}
    obstacles = len(sys.argv) > 1 and sys.argv[1] == 'obstacles'
    for name, (kw, ratio) in (OBSTACLE_CONFIGS if obstacles else CONFIGS).items():
        args = H.make_args(num_agents=N, num_landmarks=L, episode_length=25, **{k: v for k, v in kw.items() if k != 'obstacle_extension'})
        env = H.make_env(args, seed=123, obstacle_extension=bool(kw.get('obstacle_extension', False)))
        total = int(args.num_env_steps) // int(args.episode_length) // int(args.n_rollout_threads)
        ep = int(round(ratio * total))
        av, lp, lh, ls, op = [], [], [], [], []
        for r in range(R):
            env.reset(ep)
            s = H.snapshot(env)
            av.append(s['agent_values']); lp.append(s['landmark_pos']); lh.append(s['landmark_heading']); ls.append(s['landmark_speed'])
            if obstacles:
                op.append(s['obstacle_pos'])
        f, freq = features(np.array(av), np.array(lp), np.array(lh), np.array(ls), kw['dynamics_type'], np.array(op) if obstacles else None)
        q = np.linspace(0.0, 1.0, NQ)
        for k, v in f.items():
            out[f'{name}__{k}'] = np.quantile(v, q).astype(np.float32)
        meta[name] = dict(args=kw, ratio=ratio, episode=ep, total_episodes=total, resets=R, freq=freq,
                          samples={k: int(v.size) for k, v in f.items()})
        print(name, {k: round(v, 4) for k, v in freq.items()}, flush=True)
    out['meta'] = np.array(json.dumps(meta))
    path = os.path.join(os.path.dirname(HERE), 'tests', 'golden', 'aux', 'reset_samples_obstacles.npz' if obstacles else 'reset_samples.npz')
    os.makedirs(os.path.dirname(path), exist_ok=True)
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
