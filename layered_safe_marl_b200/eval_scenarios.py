"""Deterministic initial conditions of the reference's evaluation scenarios, as injectable state dicts.

The reference keeps its hand-built evaluation scenarios in a separate scenario class
(multiagent/custom_scenarios/navigation_graph_safe_eval.py); only their closed-form initial states are
reproduced here, for `B200GraphVecEnv.set_state` followed by
`reset_from_state`:

    circular                  scenario_circular_config                   navigation_graph_safe_eval.py:100-121
    two_vehicle_conflict      scenario_two_vehicle_conflicting_example   navigation_graph_safe_eval.py:383-431
    three_vehicle_conflict    scenario_three_vehicle_conflicting_example navigation_graph_safe_eval.py:320-381

The training scenario this package implements needs at least two goals per agent
(creat_relative_heading_list_from_goal_position_list asserts it, utils.py:31), while the conflict examples
carry one landmark per agent. The extra goals are appended further along the same heading, so the first
leg of every trajectory is the reference's; this is stated in the returned dict under 'note'.
"""
from __future__ import annotations

import math

import numpy as np

from .config import AirTaxiConfig, DoubleIntegratorConfig

SCENARIOS = ('circular', 'two_vehicle_conflict', 'three_vehicle_conflict')


def _blank_state(n_envs: int, N: int, L: int) -> dict:
    M = N * L
    s = {
        'agent_values': np.zeros((n_envs, N, 4)),
        'p_dist': np.zeros((n_envs, N)), 'state_time': np.zeros((n_envs, N)),
        'done': np.zeros((n_envs, N), dtype=bool), 'safety_filtered': np.zeros((n_envs, N), dtype=bool),
        'deconflicting_agent_index': -np.ones((n_envs, N), dtype=np.int32),
        'min_relative_distance': np.full((n_envs, N), np.inf), 'goal_min_time': np.full((n_envs, N), np.inf),
        'action_diff': np.zeros((n_envs, N)), 'reached_goal': np.zeros((n_envs, N), dtype=np.int32),
        'landmark_pos': np.zeros((n_envs, M, 2)), 'landmark_heading': np.zeros((n_envs, M)),
        'landmark_speed': np.zeros((n_envs, M)),
        'times_required': -np.ones((n_envs, N)), 'dists_to_goal': -np.ones((n_envs, N)),
        'dist_left_to_goal': -np.ones((n_envs, N)), 'num_agent_collisions': np.zeros((n_envs, N)),
        'current_step': np.zeros((n_envs,), dtype=np.int32), 'curriculum_ratio': np.ones((n_envs,)),
        'ep_travel_length': np.zeros((n_envs, N)), 'ep_travel_distance': np.zeros((n_envs, N)),
        'ep_done': np.zeros((n_envs, N)), 'ep_conflict': np.zeros((n_envs, N)),
        'ep_multi_engagement': np.zeros((n_envs, N)), 'ep_min_distance': np.full((n_envs, N), np.inf),
    }
    return s


def _set_landmark(s, N, order, agent, pos, heading, speed):
    m = order * N + agent           # landmark order (utils.py:10-25): index = order * N + agent
    s['landmark_pos'][:, m] = pos
    s['landmark_heading'][:, m] = heading
    s['landmark_speed'][:, m] = speed


def circular(num_agents: int, num_landmarks: int = 2, world_size: float = 4.0, dynamics_type: str = 'double_integrator',
             n_envs: int = 1) -> dict:
    """Agents on a circle heading inward, first goal at the opposite point heading outward-through,
    later goals alternate between the start point and the opposite point (the commented-out second and
    third landmark groups of scenario_circular_config)."""
    N, L = num_agents, num_landmarks
    cfg = DoubleIntegratorConfig if dynamics_type == 'double_integrator' else AirTaxiConfig
    s = _blank_state(n_envs, N, L)
    theta = np.linspace(0.0, 2.0 * np.pi, N, endpoint=False)
    radius = 0.92 * world_size / 2.0
    goal_speed = 0.5 * (cfg.V_NOMINAL + cfg.V_MIN)
    for i in range(N):
        p = np.array([radius * math.cos(theta[i]), radius * math.sin(theta[i])])
        heading_in = theta[i] + np.pi
        if dynamics_type == 'double_integrator':
            # DoubleIntegratorXYState.reset_velocity(theta) leaves the speed at its initial 0 (core.py:183-189)
            s['agent_values'][:, i] = [p[0], p[1], 0.0, 0.0]
        else:
            s['agent_values'][:, i] = [p[0], p[1], heading_in, cfg.V_MIN]   # reset_velocity(theta): speed = min_speed (core.py:137-145)
        for order in range(L):
            at_opposite = order % 2 == 0
            pos = -p if at_opposite else p
            heading = heading_in if at_opposite else theta[i]
            _set_landmark(s, N, order, i, pos, heading, goal_speed)
    s['note'] = 'circular: goals alternate opposite point / start point'
    return s


def _conflict(agents, landmark_distance, num_landmarks, n_envs):
    """agents: list of (pos, heading, speed). AirTaxi only (the reference asserts it)."""
    N, L = len(agents), num_landmarks
    s = _blank_state(n_envs, N, L)
    for i, (pos, heading, speed) in enumerate(agents):
        pos = np.asarray(pos, dtype=np.float64)
        s['agent_values'][:, i] = [pos[0], pos[1], heading, speed]
        d = np.array([math.cos(heading), math.sin(heading)])
        for order in range(L):
            _set_landmark(s, N, order, i, pos + d * landmark_distance * (order + 1), heading, AirTaxiConfig.V_NOMINAL)
    s['note'] = 'goal 0 is the reference landmark; goals 1.. continue along the same heading'
    return s


def two_vehicle_conflict(num_landmarks: int = 2, n_envs: int = 1) -> dict:
    v_nom = AirTaxiConfig.V_NOMINAL
    agents = [((0.4, 0.0), 0.0, v_nom), ((1.7, 0.3), 4.0 * np.pi / 3.0, v_nom)]
    return _conflict(agents, 3.5, num_landmarks, n_envs)


def three_vehicle_conflict(num_landmarks: int = 2, n_envs: int = 1) -> dict:
    v_nom = AirTaxiConfig.V_NOMINAL
    agents = [((0.4, 0.0), 0.0, v_nom), ((1.7, 0.3), 4.0 * np.pi / 3.0, v_nom), ((1.6, -0.6), -np.pi, AirTaxiConfig.V_MIN)]
    return _conflict(agents, 4.0, num_landmarks, n_envs)


def build(name: str, **kw) -> dict:
    if name == 'circular':
        return circular(**kw)
    if name == 'two_vehicle_conflict':
        return two_vehicle_conflict(**kw)
    if name == 'three_vehicle_conflict':
        return three_vehicle_conflict(**kw)
    raise KeyError(f"unknown eval scenario {name!r}; have {SCENARIOS}")
