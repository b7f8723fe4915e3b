"""ctypes front-end of the CPU oracle (oracle/liblsm_oracle.so). TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs import
this module. It shares NO code with the product package; the field indices below are an
independent copy of oracle/lsm_oracle.h (a CPU test checks they agree with the product's).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'liblsm_oracle.so')

AF = dict(X=0, Y=1, S2=2, S3=3, P_DIST=4, STATE_TIME=5, MIN_REL_DIST=6, GOAL_MIN_TIME=7,
          TIMES_REQ_A=8, TIMES_REQ_B=9, DISTS_GOAL_A=10, DISTS_GOAL_B=11, DIST_LEFT=12,
          EP_TRAVEL_DIST=13, EP_MIN_DIST=14, ACTION_DIFF=15)
AF_COUNT = 16
AI = dict(REACHED=0, DONE=1, SAFETY_FILTERED=2, DECONFLICT_IDX=3, NUM_COLLISIONS=4,
          EP_TRAVEL_LEN=5, EP_CONFLICT=6, EP_MULTI=7, EP_DONE=8, NUM_OBST_COLLISIONS=9)
AI_COUNT = 10
LF = dict(X=0, Y=1, HEADING=2, SPEED=3, SIN=4, COS=5)
LF_COUNT = 6
EF_COUNT = 1
EI = dict(CURRENT_STEP=0, RESET_COUNT=1, PARITY=2, JUST_RESET=3)
EI_COUNT = 4
EP_COUNT = 8


class Params(C.Structure):
    _fields_ = [('dynamics', C.c_int32), ('num_agents', C.c_int32), ('num_landmarks', C.c_int32),
                ('episode_length', C.c_int32), ('num_total_episode', C.c_int32),
                ('num_internal_step', C.c_int32), ('flags', C.c_uint32), ('num_obstacles', C.c_int32),
                ('world_size', C.c_double), ('dt', C.c_double), ('coordination_range', C.c_double),
                ('dist_thresh', C.c_double), ('heading_thresh', C.c_double), ('speed_thresh', C.c_double),
                ('goal_speed_min', C.c_double), ('goal_speed_max', C.c_double),
                ('separation_distance_target', C.c_double), ('engagement_distance_ref', C.c_double),
                ('engagement_ref_separation', C.c_double), ('cbf_rate', C.c_double),
                ('agent_max_speed', C.c_double), ('goal_rew', C.c_double),
                ('safety_violation_rew', C.c_double), ('hj_value_rew', C.c_double),
                ('potential_conflict_rew', C.c_double), ('diff_from_filtered_action_rew', C.c_double),
                ('min_reward', C.c_double), ('max_reward', C.c_double),
                ('act_tab0', C.c_double * 5), ('act_tab1', C.c_double * 5)]


class Grid(C.Structure):
    _fields_ = [('ndim', C.c_int32), ('shape', C.c_int32 * 5), ('periodic', C.c_int32 * 5),
                ('_pad', C.c_int32), ('lo', C.c_double * 5), ('hi', C.c_double * 5),
                ('separation_distance', C.c_double), ('ttr_max', C.c_double),
                ('values', C.c_void_p), ('grads', C.c_void_p)]


class Buffers(C.Structure):
    _fields_ = [('num_envs', C.c_int64), ('env_id_base', C.c_int64),
                ('agent_f64', C.c_void_p), ('agent_i32', C.c_void_p), ('landmarks', C.c_void_p),
                ('env_f64', C.c_void_p), ('env_i32', C.c_void_p),
                ('obs', C.c_void_p), ('node_obs', C.c_void_p), ('adj', C.c_void_p),
                ('reward', C.c_void_p), ('done', C.c_void_p), ('safe_action', C.c_void_p),
                ('ep_info', C.c_void_p), ('obstacles', C.c_void_p)]


def build(force=False):
    if force or not os.path.exists(LIB_PATH) or \
            os.path.getmtime(LIB_PATH) < max(os.path.getmtime(os.path.join(HERE, f))
                                             for f in ('lsm_oracle.c', 'lsm_oracle.h', '../include/lsm_math.h')):
        subprocess.check_call(['make', '-C', HERE, '-s'])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        _lib.lsmo_step.restype = C.c_int
        _lib.lsmo_reset.restype = C.c_int
        _lib.lsmo_observe.restype = C.c_int
        _lib.lsmo_interpolate.restype = C.c_double
        _lib.lsmo_magnetic_heading.restype = C.c_double
        _lib.lsmo_magnetic_heading.argtypes = [C.c_double, C.c_double, C.c_double]
        _lib.lsmo_math_eval.restype = None
        _lib.lsmo_math_eval.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        _lib.lsmo_set_relative_state_form.argtypes = [C.c_int]
        _lib.lsmo_set_interp_float32.argtypes = [C.c_int]
        _lib.lsmo_set_interp_float32.restype = None
        _lib.lsmo_get_relative_state_form.restype = C.c_int
    return _lib


def math_eval(op, a, b=None):
    """include/lsm_math.h on the host: op 0 sin(a), 1 cos(a), 2 atan2(a, b) - the float64 trigonometry the oracle and
    the CUDA kernels share (numpy's libm in the reference)."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    out = np.empty_like(a)
    bp = None
    if b is not None:
        b = np.ascontiguousarray(b, dtype=np.float64)
        bp = b.ctypes.data_as(C.c_void_p)
    lib().lsmo_math_eval(int(op), a.ctypes.data_as(C.c_void_p), bp, out.ctypes.data_as(C.c_void_p), a.size)
    return out


def set_relative_state_form(form: int):
    """0 = the reference's literal airtaxi relative position (safety_filter.py:277-284), 1 = the rotation form the
    specialised CUDA pipeline evaluates (see lsm_oracle.c). Process-global; returns the previous setting."""
    prev = lib().lsmo_get_relative_state_form()
    lib().lsmo_set_relative_state_form(int(form))
    return prev


def action_tables(dynamics):
    """np.linspace tables of MultiAgentBaseEnv._set_action (multiagent/environment.py:387-410)."""
    if dynamics == 0:
        return np.linspace(-0.5, 0.5, 5), np.linspace(-0.5, 0.5, 5)
    return np.linspace(-0.1, 0.1, 5), np.linspace(-0.001, 0.002, 5)


def make_params(pd: dict) -> Params:
    p = Params()
    for name, _ in Params._fields_:
        if name in ('_pad', 'act_tab0', 'act_tab1'):
            continue
        setattr(p, name, pd.get(name, 0) if name == 'num_obstacles' else pd[name])
    t0, t1 = action_tables(int(pd['dynamics']))
    p.act_tab0 = (C.c_double * 5)(*[float(v) for v in t0])
    p.act_tab1 = (C.c_double * 5)(*[float(v) for v in t1])
    return p


def make_grid(g):
    """g: object with lo, hi, shape, periodic, values (float32), grads (float32|None),
    separation_distance, ttr_max. Returns (Grid struct, keep-alive arrays)."""
    if g is None:
        return None, ()
    s = Grid()
    nd = len(g.shape)
    s.ndim = nd
    for d in range(nd):
        s.shape[d] = int(g.shape[d])
        s.periodic[d] = int(bool(g.periodic[d]))
        s.lo[d] = float(g.lo[d])
        s.hi[d] = float(g.hi[d])
    s.separation_distance = float(getattr(g, 'separation_distance', 0.0))
    s.ttr_max = float(getattr(g, 'ttr_max', 0.0))
    values = np.ascontiguousarray(g.values, dtype=np.float32)
    s.values = values.ctypes.data
    keep = [values]
    if getattr(g, 'grads', None) is not None:
        grads = np.ascontiguousarray(g.grads, dtype=np.float32)
        s.grads = grads.ctypes.data
        keep.append(grads)
    return s, tuple(keep)


class OracleEnv(object):
    """Batched CPU oracle with the same SoA state layout as the CUDA environment."""

    def __init__(self, params: dict, num_envs: int, value_grid=None, ttr_grid=None, seed=0,
                 env_id_base=0, nthreads=1):
        self.pd = dict(params)
        self.p = make_params(self.pd)
        self.n = int(num_envs)
        self.N = int(self.pd['num_agents'])
        self.L = int(self.pd['num_landmarks'])
        self.M = self.N * self.L
        self.O = int(self.pd.get('num_obstacles', 0))      # declared extension (lsm_oracle.h)
        self.E = self.N + self.M + self.O
        self.dyn = int(self.pd['dynamics'])
        self.D = 7 if self.dyn == 0 else 6
        self.F = 7 if (int(params["flags"]) & (1 << 9)) else (10 if self.dyn == 0 else 11)   # 7: global node features
        self.seed = int(seed)
        self.nthreads = int(nthreads)
        self.vg, self._vg_keep = make_grid(value_grid)
        self.tg, self._tg_keep = make_grid(ttr_grid)
        n, N, M, E = self.n, self.N, self.M, self.E
        self.agent_f64 = np.zeros((AF_COUNT, n, N), dtype=np.float64)
        self.agent_f64[AF['EP_MIN_DIST']] = np.inf
        self.agent_f64[AF['MIN_REL_DIST']] = np.inf
        self.agent_f64[AF['GOAL_MIN_TIME']] = np.inf
        for k in ('TIMES_REQ_A', 'TIMES_REQ_B', 'DISTS_GOAL_A', 'DISTS_GOAL_B', 'DIST_LEFT'):
            self.agent_f64[AF[k]] = -1.0
        self.agent_i32 = np.zeros((AI_COUNT, n, N), dtype=np.int32)
        self.agent_i32[AI['DECONFLICT_IDX']] = -1
        self.landmarks = np.zeros((LF_COUNT, n, M), dtype=np.float64)
        self.obstacles = np.zeros((2, n, self.O), dtype=np.float64)
        self.env_f64 = np.zeros((EF_COUNT, n), dtype=np.float64)
        self.env_i32 = np.zeros((EI_COUNT, n), dtype=np.int32)
        self.obs = np.zeros((n, N, self.D), dtype=np.float32)
        self.node_obs = np.zeros((n, N, E, self.F), dtype=np.float32)
        self.adj = np.zeros((n, N, E, E), dtype=np.float32)
        self.reward = np.zeros((n, N), dtype=np.float32)
        self.done = np.zeros((n, N), dtype=np.uint8)
        self.safe_action = np.zeros((n, N, 2), dtype=np.float64)
        self.ep_info = np.zeros((n, EP_COUNT), dtype=np.float64)
        b = Buffers()
        b.num_envs = n
        b.env_id_base = int(env_id_base)
        for name in ('agent_f64', 'agent_i32', 'landmarks', 'env_f64', 'env_i32', 'obs', 'node_obs', 'adj',
                     'reward', 'done', 'safe_action', 'ep_info'):
            setattr(b, name, getattr(self, name).ctypes.data)
        if self.O > 0:
            b.obstacles = self.obstacles.ctypes.data
        self.b = b

    def _gp(self, g):
        return C.byref(g) if g is not None else None

    def step(self, action_idx, episode=0, auto_reset=True):
        a = np.ascontiguousarray(action_idx, dtype=np.int32).reshape(self.n, self.N)
        rc = lib().lsmo_step(C.byref(self.p), self._gp(self.vg), self._gp(self.tg), C.byref(self.b),
                             C.c_void_p(a.ctypes.data), C.c_int64(int(episode)), C.c_uint64(self.seed),
                             C.c_int(int(auto_reset)), C.c_int(self.nthreads))
        if rc != 0:
            raise RuntimeError(f"lsmo_step failed rc={rc}")

    def reset(self, episode=0, env_mask=None, sample=True):
        m = None
        if env_mask is not None:
            m = np.ascontiguousarray(env_mask, dtype=np.uint8)
        rc = lib().lsmo_reset(C.byref(self.p), self._gp(self.vg), self._gp(self.tg), C.byref(self.b),
                              C.c_void_p(m.ctypes.data) if m is not None else None,
                              C.c_int64(int(episode)), C.c_uint64(self.seed), C.c_int(int(sample)),
                              C.c_int(self.nthreads))
        if rc != 0:
            raise RuntimeError(f"lsmo_reset failed rc={rc}")

    def observe(self):
        rc = lib().lsmo_observe(C.byref(self.p), C.byref(self.b), C.c_int(self.nthreads))
        if rc != 0:
            raise RuntimeError(f"lsmo_observe failed rc={rc}")

    # ---- named-state interchange (same keys as oracle/ref_harness.snapshot) -------------------
    def set_state(self, s: dict, env=None):
        """s: named arrays with a leading env axis (n, ...) or, with `env=e`, for one env."""
        sel = slice(None) if env is None else env
        f, i = self.agent_f64, self.agent_i32
        v = np.asarray(s['agent_values'], dtype=np.float64)
        f[AF['X']][sel] = v[..., 0]; f[AF['Y']][sel] = v[..., 1]
        f[AF['S2']][sel] = v[..., 2]; f[AF['S3']][sel] = v[..., 3]
        f[AF['P_DIST']][sel] = s['p_dist']; f[AF['STATE_TIME']][sel] = s['state_time']
        f[AF['MIN_REL_DIST']][sel] = s['min_relative_distance']
        f[AF['GOAL_MIN_TIME']][sel] = s['goal_min_time']
        for k in ('TIMES_REQ_A', 'TIMES_REQ_B'):
            f[AF[k]][sel] = s['times_required']
        for k in ('DISTS_GOAL_A', 'DISTS_GOAL_B'):
            f[AF[k]][sel] = s['dists_to_goal']
        f[AF['DIST_LEFT']][sel] = s['dist_left_to_goal']
        f[AF['EP_TRAVEL_DIST']][sel] = s['ep_travel_distance']
        f[AF['EP_MIN_DIST']][sel] = s['ep_min_distance']
        f[AF['ACTION_DIFF']][sel] = s['action_diff']
        i[AI['REACHED']][sel] = s['reached_goal']; i[AI['DONE']][sel] = s['done']
        i[AI['SAFETY_FILTERED']][sel] = s['safety_filtered']
        i[AI['DECONFLICT_IDX']][sel] = s['deconflicting_agent_index']
        i[AI['NUM_COLLISIONS']][sel] = np.asarray(s['num_agent_collisions']).astype(np.int32)
        i[AI['EP_TRAVEL_LEN']][sel] = np.asarray(s['ep_travel_length']).astype(np.int32)
        i[AI['EP_CONFLICT']][sel] = np.asarray(s['ep_conflict']).astype(np.int32)
        i[AI['EP_MULTI']][sel] = np.asarray(s['ep_multi_engagement']).astype(np.int32)
        i[AI['EP_DONE']][sel] = np.asarray(s['ep_done']).astype(np.int32)
        lp = np.asarray(s['landmark_pos'], dtype=np.float64)
        lh = np.asarray(s['landmark_heading'], dtype=np.float64)
        self.landmarks[LF['X']][sel] = lp[..., 0]; self.landmarks[LF['Y']][sel] = lp[..., 1]
        self.landmarks[LF['HEADING']][sel] = lh
        self.landmarks[LF['SPEED']][sel] = s['landmark_speed']
        self.landmarks[LF['SIN']][sel] = math_eval(0, lh); self.landmarks[LF['COS']][sel] = math_eval(1, lh)
        self.env_f64[0][sel] = s['curriculum_ratio']
        self.env_i32[EI['CURRENT_STEP']][sel] = s['current_step']
        if self.O > 0:
            op = np.asarray(s['obstacle_pos'], dtype=np.float64)
            self.obstacles[0][sel] = op[..., 0]; self.obstacles[1][sel] = op[..., 1]
            i[AI['NUM_OBST_COLLISIONS']][sel] = np.asarray(s['num_obstacle_collisions']).astype(np.int32)

    def get_state(self):
        f, i = self.agent_f64, self.agent_i32
        par = self.env_i32[EI['PARITY']][:, None].astype(bool)
        s = {}
        s['agent_values'] = np.stack([f[AF['X']], f[AF['Y']], f[AF['S2']], f[AF['S3']]], axis=-1)
        s['p_dist'] = f[AF['P_DIST']].copy(); s['state_time'] = f[AF['STATE_TIME']].copy()
        s['min_relative_distance'] = f[AF['MIN_REL_DIST']].copy()
        s['goal_min_time'] = f[AF['GOAL_MIN_TIME']].copy()
        s['times_required'] = np.where(par, f[AF['TIMES_REQ_B']], f[AF['TIMES_REQ_A']])
        s['dists_to_goal'] = np.where(par, f[AF['DISTS_GOAL_B']], f[AF['DISTS_GOAL_A']])
        s['dist_left_to_goal'] = f[AF['DIST_LEFT']].copy()
        s['ep_travel_distance'] = f[AF['EP_TRAVEL_DIST']].copy()
        s['ep_min_distance'] = f[AF['EP_MIN_DIST']].copy()
        s['action_diff'] = f[AF['ACTION_DIFF']].copy()
        s['reached_goal'] = i[AI['REACHED']].copy(); s['done'] = i[AI['DONE']].astype(bool)
        s['safety_filtered'] = i[AI['SAFETY_FILTERED']].astype(bool)
        s['deconflicting_agent_index'] = i[AI['DECONFLICT_IDX']].copy()
        s['num_agent_collisions'] = i[AI['NUM_COLLISIONS']].astype(np.float64)
        s['ep_travel_length'] = i[AI['EP_TRAVEL_LEN']].astype(np.float64)
        s['ep_conflict'] = i[AI['EP_CONFLICT']].astype(np.float64)
        s['ep_multi_engagement'] = i[AI['EP_MULTI']].astype(np.float64)
        s['ep_done'] = i[AI['EP_DONE']].astype(np.float64)
        s['landmark_pos'] = np.stack([self.landmarks[LF['X']], self.landmarks[LF['Y']]], axis=-1)
        s['landmark_heading'] = self.landmarks[LF['HEADING']].copy()
        s['landmark_speed'] = self.landmarks[LF['SPEED']].copy()
        s['curriculum_ratio'] = self.env_f64[0].copy()
        s['current_step'] = self.env_i32[EI['CURRENT_STEP']].copy()
        if self.O > 0:
            s['obstacle_pos'] = np.stack([self.obstacles[0], self.obstacles[1]], axis=-1)
            s['num_obstacle_collisions'] = i[AI['NUM_OBST_COLLISIONS']].astype(np.float64)
        return s
