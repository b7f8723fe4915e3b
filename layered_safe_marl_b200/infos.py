"""Lazily materialised `infos` of the vectorised environment.

The reference returns, per env, a list of N python dicts (navigation_graph_safe.py:386-450 +
`individual_reward`, environment.py:1024-1028) and, on the step where the env auto-reset, an
(N+1)-th dict with the 8-key episode summary (onpolicy/envs/env_wrappers.py:865-874). Building
those for thousands of envs every step would dominate the step, so `LazyInfos` is a sequence proxy
over the device state: nothing is copied until an element is read, then ONE device->host copy
serves every env. A proxy is valid until the next `step` / `reset` of its environment.

`compute_agent_infos` is pure numpy (no CUDA) so it is unit-testable on CPU.
"""
from __future__ import annotations

import numpy as np

from . import layout as LY

AGENT_INFO_KEYS = ('individual_reward', 'id', 'position', 'min_relative_distance', 'Dist_to_goal',
                   'Time_req_to_goal', 'Num_agent_collisions', 'Num_obst_collisions', 'Distance_mean',
                   'Distance_variance', 'Mean_by_variance', 'Dists_traveled', 'Time_taken', 'Time_mean',
                   'Time_stddev', 'Time_mean_by_stddev', 'Min_time_to_goal', 'Departed', 'Safety filtered',
                   'Safety violated')


def _mixture_mean_std(new: np.ndarray, old: np.ndarray):
    """Agent i's info dict is built inside the sequential per-agent loop (environment.py:979-1029), so
    the world arrays it averages hold THIS step's value for agents <= i and LAST step's for agents > i.
    new/old: (n, N) -> mean, std of shape (n, N) [index i]."""
    n, N = new.shape
    k = np.arange(N)
    use_new = k[None, :] <= k[:, None]                                   # (i, k)
    mix = np.where(use_new[None], new[:, None, :], old[:, None, :])      # (n, i, k)
    mean = mix.mean(axis=2)
    std = mix.std(axis=2)
    return mean, std


def compute_agent_infos(agent_f64: np.ndarray, agent_i32: np.ndarray, env_i32: np.ndarray,
                        individual_reward: np.ndarray, separation_distance: np.ndarray) -> dict:
    """-> {key: (n, N[, 2]) array}. `separation_distance`: (n,) scenario value of each env's curriculum."""
    f, i = agent_f64, agent_i32
    par = env_i32[LY.EI_PARITY][:, None].astype(bool)
    times_new = np.where(par, f[LY.AF_TIMES_REQ_B], f[LY.AF_TIMES_REQ_A])
    times_old = np.where(par, f[LY.AF_TIMES_REQ_A], f[LY.AF_TIMES_REQ_B])
    dists_new = np.where(par, f[LY.AF_DISTS_GOAL_B], f[LY.AF_DISTS_GOAL_A])
    dists_old = np.where(par, f[LY.AF_DISTS_GOAL_A], f[LY.AF_DISTS_GOAL_B])
    d_mean, d_std = _mixture_mean_std(dists_new, dists_old)
    t_mean, t_std = _mixture_mean_std(times_new, times_old)
    n, N = times_new.shape
    out = {
        'individual_reward': np.asarray(individual_reward, dtype=np.float64),
        'id': np.broadcast_to(np.arange(N, dtype=np.int64), (n, N)),
        'position': np.stack([f[LY.AF_X], f[LY.AF_Y]], axis=-1),
        'min_relative_distance': f[LY.AF_MIN_REL_DIST],
        'Dist_to_goal': f[LY.AF_DIST_LEFT],
        'Time_req_to_goal': times_new,
        'Num_agent_collisions': i[LY.AI_NUM_COLLISIONS].astype(np.float64),
        'Num_obst_collisions': i[LY.AI_NUM_OBST_COLLISIONS].astype(np.float64),   # 0 unless the obstacle extension is on
        'Distance_mean': d_mean,
        'Distance_variance': d_std,
        'Mean_by_variance': d_mean / (d_std + 0.0001),
        'Dists_traveled': dists_new,
        'Time_taken': times_new,
        'Time_mean': t_mean,
        'Time_stddev': t_std,
        'Time_mean_by_stddev': t_mean / (t_std + 0.0001),
        'Min_time_to_goal': f[LY.AF_GOAL_MIN_TIME],
        'Departed': np.ones((n, N), dtype=bool),
        'Safety filtered': i[LY.AI_SAFETY_FILTERED].astype(bool),
        'Safety violated': f[LY.AF_MIN_REL_DIST] < np.asarray(separation_distance)[:, None],
    }
    return out


def apply_terminal_snapshot(agent_f64, agent_i32, env_i32, ratio, term_f64, term_i32, term_ratio):
    """For the envs that auto-reset in this step the device state already belongs to the NEW episode, while the
    reference's graphworker returns the TERMINAL step's info dicts and only appends the episode summary
    (onpolicy/envs/env_wrappers.py:861-874). The kernels snapshot the fields `info_callback` reads just before the
    reset (include/lsm_b200.h LSM_TF_* / LSM_TI_*); this puts them back, in place, for those envs (host copies)."""
    jr = env_i32[LY.EI_JUST_RESET].astype(bool)
    if not jr.any():
        return ratio
    f, i = agent_f64, agent_i32
    par = env_i32[LY.EI_PARITY].astype(bool)
    for dst, src in ((LY.AF_X, LY.TF_X), (LY.AF_Y, LY.TF_Y), (LY.AF_MIN_REL_DIST, LY.TF_MIN_REL_DIST),
                     (LY.AF_DIST_LEFT, LY.TF_DIST_LEFT), (LY.AF_GOAL_MIN_TIME, LY.TF_GOAL_MIN_TIME)):
        f[dst][jr] = term_f64[src][jr]
    new_b, new_a = jr & par, jr & ~par          # the slot `parity` names holds the newest values
    for slot_a, slot_b, t_new, t_old in ((LY.AF_TIMES_REQ_A, LY.AF_TIMES_REQ_B, LY.TF_TIMES_REQ_NEW, LY.TF_TIMES_REQ_OLD),
                                         (LY.AF_DISTS_GOAL_A, LY.AF_DISTS_GOAL_B, LY.TF_DISTS_GOAL_NEW, LY.TF_DISTS_GOAL_OLD)):
        f[slot_b][new_b] = term_f64[t_new][new_b]; f[slot_a][new_b] = term_f64[t_old][new_b]
        f[slot_a][new_a] = term_f64[t_new][new_a]; f[slot_b][new_a] = term_f64[t_old][new_a]
    i[LY.AI_NUM_COLLISIONS][jr] = term_i32[LY.TI_NUM_COLLISIONS][jr]
    i[LY.AI_SAFETY_FILTERED][jr] = term_i32[LY.TI_SAFETY_FILTERED][jr]
    i[LY.AI_NUM_OBST_COLLISIONS][jr] = term_i32[LY.TI_NUM_OBST_COLLISIONS][jr]
    ratio = np.array(ratio, dtype=np.float64, copy=True)
    ratio[jr] = term_ratio[jr]
    return ratio


def separation_distance_of(params, ratio: np.ndarray) -> np.ndarray:
    """scenario.separation_distance from the per-env curriculum ratio (navigation_graph_safe.py:349-363)."""
    from .config import FLAG_SEPARATION_DISTANCE_CURRICULUM
    ratio = np.asarray(ratio, dtype=np.float64)
    start, end, num_steps = 0.2, 0.75, 4
    cont = (num_steps - 1) * np.clip(ratio - start, 0, end - start) / (end - start)
    stair = (1 + np.floor(cont)) / num_steps
    stair = np.where(ratio < start, 0.0, np.where(ratio > end, 1.0, stair))
    sep_ratio = 1 - np.cos(stair * 0.5 * np.pi)
    target = params.separation_distance_target
    init = 0.0 if (params.flags & FLAG_SEPARATION_DISTANCE_CURRICULUM) else target
    return init * (1.0 - sep_ratio) + target * sep_ratio


class LazyInfos:
    """Sequence over envs; `infos[e]` is a list of N dicts (+ the episode-summary dict if env e just reset)."""

    def __init__(self, env, step_id: int, reset_only: bool):
        self._env = env
        self._step_id = step_id
        self._reset_only = reset_only
        self._host = None

    def _check_live(self):
        if self._env._step_id != self._step_id:
            raise RuntimeError("stale infos: the environment has stepped since these infos were returned "
                               "(read them before the next step, like the returned tensors)")

    def _materialise(self):
        if self._host is not None:
            return self._host
        self._check_live()
        env = self._env
        h = {}
        h['ep_info'] = env.ep_info.cpu().numpy()
        h['env_i32'] = env.env_i32.cpu().numpy()
        if not self._reset_only:
            af = env.agent_f64.cpu().numpy()
            ai = env.agent_i32.cpu().numpy()
            ratio = env.env_f64[LY.EF_CURRICULUM_RATIO].cpu().numpy()
            if h['env_i32'][LY.EI_JUST_RESET].any():
                ratio = apply_terminal_snapshot(af, ai, h['env_i32'], ratio, env.term_f64.cpu().numpy(),
                                                env.term_i32.cpu().numpy(), env.term_env_f64.cpu().numpy())
            rew = (env.reward_individual if env.reward_individual is not None else env.reward).cpu().numpy()
            h['agent'] = compute_agent_infos(af, ai, h['env_i32'], rew, separation_distance_of(env.params, ratio))
        self._host = h
        return h

    def just_reset(self) -> np.ndarray:
        return self._materialise()['env_i32'][LY.EI_JUST_RESET].astype(bool)

    def episode_info(self, e: int) -> dict:
        ep = self._materialise()['ep_info'][e]
        return {k: float(ep[j]) for j, k in enumerate(LY.EP_INFO_KEYS)}

    def __len__(self):
        return self._env.num_envs

    def __iter__(self):
        for e in range(len(self)):
            yield self[e]

    def __getitem__(self, e):
        h = self._materialise()
        if e < 0:
            e += len(self)
        if not 0 <= e < len(self):
            raise IndexError(e)
        if self._reset_only:
            return self.episode_info(e)        # reset() returns one summary dict per env
        a = h['agent']
        N = self._env.N
        out = []
        for i in range(N):
            d = {}
            for k in AGENT_INFO_KEYS:
                v = a[k][e, i]
                d[k] = v.copy() if isinstance(v, np.ndarray) and v.ndim else (v.item() if hasattr(v, 'item') else v)
            out.append(d)
        if h['env_i32'][LY.EI_JUST_RESET][e]:
            out.append(self.episode_info(e))
        return out
