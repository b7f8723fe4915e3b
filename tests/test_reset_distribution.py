"""the reset sampler against the reference's `random_scenario`, DISTRIBUTIONALLY.

The reference draws from numpy's global MT19937 stream with data-dependent rejection loops
(navigation_graph_safe.py:1199-1367, utils.py:39-68); the device sampler is a counter-based Philox stream with the same
draw order, so the two cannot agree sample by sample. tests/golden/aux/reset_samples.npz holds 1 024 quantiles of every
scalar feature of 2 500 reference resets per configuration (oracle/gen_reset_samples.py, double integrator and airtaxi,
curriculum ratio 0 / 0.5 / 1, filter argument on / off). Here the same features of 5 000 resets of the sampler under
test are compared with a two-sample Kolmogorov-Smirnov statistic (alpha = 1e-3) and the event frequencies (goal copied
from the previous agent, fixed speed pattern, unperturbed last heading) with a 5-sigma binomial bound.

CPU: the C oracle's sampler. GPU: the CUDA sampler (which the parity tests also show to be bit-identical to the oracle's).
"""
import json
import os
import sys

import numpy as np
import pytest

import _golden as G

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, 'oracle'))
from gen_reset_samples import features, N, L, NQ  # noqa: E402  (pure numpy; the generator's own feature definitions)

FIX, META = {}, {}
for _name in ('reset_samples.npz', 'reset_samples_obstacles.npz'):      # the second: declared obstacle extension
    _z = np.load(os.path.join(REPO, 'tests', 'golden', 'aux', _name))
    FIX.update({k: _z[k] for k in _z.files if k != 'meta'})
    META.update(json.loads(str(_z['meta'])))
N_RESETS = 5000


def ks_against_quantiles(samples, quantiles):
    """sup |F_samples - F_ref| with F_ref the empirical CDF of the stored quantile points; atoms handled on both sides."""
    s = np.sort(np.asarray(samples, dtype=np.float64))
    q = np.sort(np.asarray(quantiles, dtype=np.float64))
    pts = np.unique(q)
    d = 0.0
    for side in ('right', 'left'):
        fs = np.searchsorted(s, pts, side=side) / s.size
        fq = np.searchsorted(q, pts, side=side) / q.size
        d = max(d, float(np.abs(fs - fq).max()))
    return d


def _check(name, state):
    m = META[name]
    dyn = m['args']['dynamics_type']
    f, freq = features(np.asarray(state['agent_values']), np.asarray(state['landmark_pos']),
                       np.asarray(state['landmark_heading']), np.asarray(state['landmark_speed']), dyn,
                       np.asarray(state['obstacle_pos']) if 'obstacle_pos' in state else None)
    if 'agent_nearest_obstacle' in f:
        assert f['agent_nearest_obstacle'].min() >= 1.05 * (0.05 + 0.05), "an accepted agent position collides with an obstacle"
    report = []
    for k, v in f.items():
        ref = FIX[f'{name}__{k}']
        n_ref = min(NQ, m['samples'][k])
        crit = 1.95 * np.sqrt(1.0 / n_ref + 1.0 / v.size) + 1.0 / NQ       # alpha = 1e-3 + quantile discretisation
        if np.ptp(ref) < 1e-12:                                             # a constant in the reference (e.g. airtaxi goal speed)
            assert np.allclose(v, ref[0], rtol=0, atol=1e-6), f"{name} {k}: reference is the constant {ref[0]}"
            continue
        d = ks_against_quantiles(v.astype(np.float32), ref)
        report.append((k, round(d, 4), round(crit, 4)))
        assert d < crit, f"{name}: feature '{k}' KS distance {d:.4f} >= {crit:.4f}"
    for k, p_ref in m['freq'].items():
        trials_ref = m['resets'] * (N - 1 if k.startswith('copy') else N)
        trials = state['agent_values'].shape[0] * (N - 1 if k.startswith('copy') else N)
        p = 0.5 * (p_ref + freq[k])
        sigma = np.sqrt(max(p * (1 - p), 1e-12) * (1.0 / trials_ref + 1.0 / trials))
        assert abs(freq[k] - p_ref) <= 5.0 * sigma + 1e-9, f"{name}: frequency '{k}' {freq[k]:.4f} vs reference {p_ref:.4f}"
        report.append((k, round(freq[k], 4), round(p_ref, 4)))
    print(name, report)


def _args(name):
    m = META[name]
    return G.default_args(num_agents=N, num_landmarks=L, episode_length=25, **m['args']), m['episode']


@pytest.mark.parametrize('name', sorted(META))
def test_oracle_reset_sampler_matches_reference_distribution(name):
    import oracle_env as O
    from layered_safe_marl_b200 import config as cfg
    args, ep = _args(name)
    params = cfg.scenario_params_from_args(args, binary_cfg=G.BinaryFlags({}))
    assert params.num_total_episode == META[name]['total_episodes']
    vg, tg = G.value_grid_for(params)
    ora = O.OracleEnv(params.asdict(), N_RESETS, value_grid=vg, ttr_grid=tg, seed=2025, nthreads=8)
    ora.reset(episode=ep, sample=True)
    _check(name, ora.get_state())


@pytest.mark.gpu
@pytest.mark.parametrize('name', sorted(META))
def test_cuda_reset_sampler_matches_reference_distribution(name):
    from layered_safe_marl_b200 import B200GraphVecEnv
    args, ep = _args(name)
    env = B200GraphVecEnv(args, num_envs=N_RESETS, seed=77)
    env.reset(ep)
    _check(name, env.get_state())
