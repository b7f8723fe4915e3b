"""Drive the UNMODIFIED reference (`/root/reference/multiagent`) offline. TEST INFRASTRUCTURE ONLY.

Only `oracle/gen_golden.py` (and ad-hoc pinning scripts) import this; it needs
`/root/reference`, which exists in the build container but NOT on the GPU box, so nothing
in `tests/ -m gpu`, `smoke()` or `bench.py` may depend on it at run time.

What it does
  * puts `oracle/ref_stubs` (functional numpy stand-ins for jax / hj_reachability /
    hj_reachability_utils / cvxpy, inert gym / pyglet / casadi) and `/root/reference` on sys.path;
  * writes the synthetic value / TTR grids (layered_safe_marl_b200.hj_grid) as pickles under a
    scratch `data/` directory and chdirs there, because the reference opens
    `data/crazyflies_value_function.pkl` etc. relative to the cwd (multiagent/config.py:29,30,62);
  * builds `GraphMPEEnv(args)` exactly like `scripts/train_mpe.py:23-45`;
  * snapshots / injects the full simulator state so the same state can be loaded into the
    C oracle and the CUDA environment.
"""
from __future__ import annotations

import argparse
import os
import pickle
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REFERENCE = os.environ.get('LSM_REFERENCE_PATH', '/root/reference')

_SETUP_DONE = False


def setup_reference():
    """Idempotent: stubs + reference on sys.path, synthetic data pickles in a scratch cwd."""
    global _SETUP_DONE
    if _SETUP_DONE:
        return
    if not os.path.isdir(os.path.join(REFERENCE, 'multiagent')):
        raise RuntimeError(f"reference not found at {REFERENCE} (expected in the build container only)")
    for p in (REFERENCE, os.path.join(HERE, 'ref_stubs'), REPO):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, REPO)
    sys.path.insert(0, REFERENCE)
    sys.path.insert(0, os.path.join(HERE, 'ref_stubs'))

    from layered_safe_marl_b200 import hj_grid as G
    from layered_safe_marl_b200.config import AirTaxiConfig, DoubleIntegratorConfig
    from hj_reachability_utils.common import GridMetaData, HjData, TtrData

    scratch = tempfile.mkdtemp(prefix='lsm_ref_')
    os.makedirs(os.path.join(scratch, 'data'))
    di = HjData(G.synthetic_di_stored_values(),
                GridMetaData(G.DI_GRID_LO, G.DI_GRID_HI, G.DI_GRID_SHAPE, ()),
                DoubleIntegratorConfig.SEPARATION_DISTANCE)
    at = HjData(G.synthetic_airtaxi_stored_values(),
                GridMetaData(G.AIRTAXI_GRID_LO, G.AIRTAXI_GRID_HI, G.AIRTAXI_GRID_SHAPE, (2,)),
                AirTaxiConfig.SEPARATION_DISTANCE)
    ttr_grid = G.synthetic_ttr_grid()
    ttr = TtrData(ttr_grid.values, GridMetaData(G.TTR_GRID_LO, G.TTR_GRID_HI, G.TTR_GRID_SHAPE, (2,)),
                  G.TTR_MAX)
    for name, obj in (('crazyflies_value_function.pkl', di), ('airtaxi_value_function.pkl', at),
                      ('airtaxi_ttr_function.pkl', ttr)):
        with open(os.path.join(scratch, 'data', name), 'wb') as f:
            pickle.dump(obj, f)
    os.chdir(scratch)
    _SETUP_DONE = True


def make_args(**kw):
    """The Namespace fields `make_world` / `GraphMPEEnv` read (train.sh:86-114 defaults)."""
    d = dict(scenario_name='navigation_graph_safe', dynamics_type='double_integrator', num_agents=3,
             num_scripted_agents=0, num_obstacles=0, collaborative=False, use_dones=False,
             episode_length=25, num_env_steps=5_000_000, n_rollout_threads=32, world_size=4,
             num_landmarks=2, use_safety_filter=False, num_internal_step=1, graph_feat_type='relative',
             use_masking=True, num_walls=0, zeroshift=3, discrete_action=True, algorithm_name='rmappo')
    d.update(kw)
    return argparse.Namespace(**d)


BINARY_FLAGS = ('SAFETY_VIOLATION', 'HJ_VALUE', 'POTENTIAL_CONFLICT', 'SEPARATION_DISTANCE_CURRICULUM',
                'INITIAL_PHASE_USE_SAFETY_FILTER', 'DIFF_FROM_FILTERED_ACTION')


def _obstacle_extension(module):
    """The DECLARED obstacle extension (BASELINE config 3 '+ obstacles'), applied to a freshly loaded scenario
    module WITHOUT touching the reference's files. With num_obstacles > 0 the reference raises in exactly two statements:
    `_get_entity_feat_relative` has no obstacle branch (navigation_graph_safe.py:1064-1065,1087) and `graph_observation`
    indexes the E x E distance matrix with a mask of N(1+L) entries (:975-989). Everything else - obstacle creation,
    placement, collision counting, distances, the 'global' features - is the reference's own code and runs unmodified.
    The two statements are completed as declared: an obstacle's relative node features are the reference's landmark builders
    (utils.py:174-199,231-255) called with heading 0 and speed 0, entity type 2; the disconnect mask is padded with False."""
    import multiagent.custom_scenarios.utils as U
    Sc = module.Scenario
    orig_rel = Sc._get_entity_feat_relative

    def _get_entity_feat_relative(self, agent, entity, world):
        if 'obstacle' in entity.name:
            if agent.dynamics_type == module.EntityDynamicsType.DoubleIntegratorXY:
                f = U.get_landmark_node_observation_relative_without_heading(entity.state.p_pos, 0.0, 0.0, agent.state)
            else:
                f = U.get_landmark_node_observation_relative_with_heading(entity.state.p_pos, 0.0, 0.0, agent.state)
            f = np.array(f, dtype=np.float64)
            f[-1] = U.entity_mapping['obstacle']
            return f
        return orig_rel(self, agent, entity, world)

    def graph_observation(self, agent, world):
        # navigation_graph_safe.py:956-994 with the mask padded for the obstacles
        node_obs = []
        for entity in world.entities:
            if world.graph_feat_type == 'global':
                node_obs.append(self._get_entity_feat_global(entity, world))
            elif world.graph_feat_type == 'relative':
                node_obs.append(self._get_entity_feat_relative(agent, entity, world))
        node_obs = np.array(node_obs)
        adj = world.cached_dist_mag
        disconnected_mask = [entity.done or not entity.departed for entity in world.agents]
        for i_landmark, _ in enumerate(world.landmarks):
            disconnected_mask.append(self.reached_goal[i_landmark % self.num_agents] > i_landmark // self.num_agents)
        disconnected_mask += [False] * len(world.obstacles)
        adj[disconnected_mask, :] = 0
        adj[:, disconnected_mask] = 0
        connect_mask = ((adj < self.max_edge_dist) & (adj > 0)).astype(np.float32)
        return node_obs, adj * connect_mask

    Sc._get_entity_feat_relative = _get_entity_feat_relative
    Sc.graph_observation = graph_observation


def make_env(args, seed=0, interp_float32=False, obstacle_extension=False, **binary_flags):
    """GraphMPEEnv(args) with the RewardBinaryConfig switches set the way the reference README says
    to set them (edit the class attributes), then env.seed(seed) (scripts/train_mpe.py:38).
    obstacle_extension: see _obstacle_extension (not reference behaviour; the default leaves the reference untouched)."""
    setup_reference()
    import hj_reachability
    hj_reachability.FLOAT32_INTERPOLATION = bool(interp_float32)      # which declared Grid.interpolate arithmetic the stub runs
    import multiagent.config as C
    for k in BINARY_FLAGS:
        setattr(C.RewardBinaryConfig, k, bool(binary_flags.get(k, False)))
    import multiagent.MPE_env as ME
    np.random.seed(seed)  # make_world itself draws (wall_length) and resets once
    if obstacle_extension:
        orig_load = ME.load     # custom_scenarios.load executes a fresh copy of the scenario module on every call

        def load(name):
            module = orig_load(name)
            _obstacle_extension(module)
            return module
        ME.load = load
        try:
            env = ME.GraphMPEEnv(args)
        finally:
            ME.load = orig_load
    else:
        env = ME.GraphMPEEnv(args)
    env.seed(seed)
    return env


def scenario_of(env):
    return env.reward_callback.__self__


def snapshot(env):
    """Everything the step path reads or carries over, as plain numpy."""
    world = env.world
    sc = scenario_of(env)
    n = len(world.agents)
    s = {}
    s['agent_values'] = np.array([a.state.values for a in world.agents], dtype=np.float64)
    s['p_dist'] = np.array([a.state.p_dist for a in world.agents], dtype=np.float64)
    s['state_time'] = np.array([a.state.time for a in world.agents], dtype=np.float64)
    s['done'] = np.array([bool(a.done) for a in world.agents])
    s['safety_filtered'] = np.array([bool(a.safety_filtered) for a in world.agents])
    s['deconflicting_agent_index'] = np.array([int(a.deconflicting_agent_index) for a in world.agents], dtype=np.int32)
    s['min_relative_distance'] = np.array([a.min_relative_distance for a in world.agents], dtype=np.float64)
    s['goal_min_time'] = np.array([a.goal_min_time for a in world.agents], dtype=np.float64)
    s['action_diff'] = np.array([a.action_diff for a in world.agents], dtype=np.float64)
    s['reached_goal'] = np.array(sc.reached_goal, dtype=np.int32)
    s['landmark_pos'] = np.array([l.state.p_pos for l in world.landmarks], dtype=np.float64)
    s['landmark_heading'] = np.array([l.heading for l in world.landmarks], dtype=np.float64)
    s['landmark_speed'] = np.array([l.speed for l in world.landmarks], dtype=np.float64)
    s['times_required'] = np.array(world.times_required, dtype=np.float64)
    s['dists_to_goal'] = np.array(world.dists_to_goal, dtype=np.float64)
    s['dist_left_to_goal'] = np.array(world.dist_left_to_goal, dtype=np.float64)
    s['num_agent_collisions'] = np.array(world.num_agent_collisions, dtype=np.float64)
    s['current_step'] = np.int32(env.current_step)
    s['curriculum_ratio'] = np.float64(sc.curriculum_ratio)
    s['world_use_safety_filter'] = np.bool_(world.use_safety_filter)
    s['separation_distance'] = np.float64(sc.separation_distance)
    s['engagement_distance'] = np.float64(sc.engagement_distance)
    # episode statistics accumulators (environment.py:886-926)
    s['ep_travel_length'] = np.array(env.episode_agent_travel_length_list, dtype=np.float64)
    s['ep_travel_distance'] = np.array(env.episode_agent_travel_distance_list, dtype=np.float64)
    s['ep_done'] = np.array(env.episode_agent_done_list, dtype=np.float64)
    s['ep_conflict'] = np.array(env.episode_agent_conflict_occurance_list, dtype=np.float64)
    s['ep_multi_engagement'] = np.array(env.episode_agent_in_multiple_engagement_list, dtype=np.float64)
    s['ep_min_distance'] = np.array(env.episode_agent_min_distance_list, dtype=np.float64)
    if len(world.obstacles) > 0:      # obstacle extension only
        s['obstacle_pos'] = np.array([o.state.p_pos for o in world.obstacles], dtype=np.float64)
        s['num_obstacle_collisions'] = np.array(world.num_obstacle_collisions, dtype=np.float64)
    assert s['agent_values'].shape == (n, 4)
    return s


def one_hot(idx, n_actions=25):
    return [np.eye(n_actions)[int(i)] for i in idx]


def info_arrays(infos, n):
    """Per-agent info dicts (navigation_graph_safe.py:425-450) -> dict of arrays."""
    keys = ['individual_reward', 'min_relative_distance', 'Dist_to_goal', 'Time_req_to_goal',
            'Num_agent_collisions', 'Num_obst_collisions', 'Distance_mean', 'Distance_variance',
            'Mean_by_variance', 'Dists_traveled', 'Time_taken', 'Time_mean', 'Time_stddev',
            'Time_mean_by_stddev', 'Min_time_to_goal', 'Departed', 'Safety filtered', 'Safety violated']
    out = {k: np.array([float(infos[i][k]) for i in range(n)], dtype=np.float64) for k in keys}
    out['position'] = np.array([np.array(infos[i]['position'], dtype=np.float64) for i in range(n)])
    out['id'] = np.array([int(infos[i]['id']) for i in range(n)], dtype=np.int32)
    return out
