"""Physical constants, reward weights and reward switches of the navigation scenario.

Host-side mirror of the reference's `multiagent/config.py:3-83` (same class and
attribute names, same values) so that code written against the reference reads the
same here. In the reference the switches of `RewardBinaryConfig` are toggled by editing
the source file (reference README.md:88-90); here they are captured into an explicit,
immutable `ScenarioParams` at environment construction and handed to the CUDA kernels
as launch constants.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field, asdict


class AirTaxiConfig:
    V_MIN = 60 * 0.514444 * 0.001  # knots -> km/s
    V_MAX = 175 * 0.514444 * 0.001
    V_NOMINAL = 110 * 0.514444 * 0.001
    ACCEL_MIN = -0.001  # km/s^2
    ACCEL_MAX = 0.002
    ANGULAR_RATE_MAX = 0.1  # rad/s
    MOTION_PRIM_ACCEL_OPTIONS = 5
    MOTION_PRIM_ANGRATE_OPTIONS = 5
    CBF_RATE = 3.0
    ENGAGEMENT_DISTANCE = 1.4
    ENGAGEMENT_DISTANCE_REFERENCE_SEPARATION_DISTANCE = 2200 * 0.0003048
    DT = 1.0
    DISTANCE_TO_GOAL_THRESHOLD = 0.35
    GOAL_HEADING_THRESHOLD = math.pi / 4
    GOAL_SPEED_THRESHOLD = 0.03
    SEPARATION_DISTANCE = 1500 * 0.0003048  # ft -> km
    COORDINATION_RANGE = 3 * 1.60934  # 3 miles -> km
    VALUE_FUNCTION_FILE_NAME = 'data/airtaxi_value_function.pkl'
    TTR_FILE_NAME = 'data/airtaxi_ttr_function.pkl'


class DoubleIntegratorConfig:
    VX_MIN = -0.5
    VX_MAX = 0.5
    VY_MIN = -0.5
    VY_MAX = 0.5
    V_MIN = 0.1
    V_NOMINAL = 0.5
    V_MAX = math.sqrt(VX_MAX ** 2 + VY_MAX ** 2)
    ACCELX_MIN = -0.5
    ACCELX_MAX = 0.5
    ACCELY_MIN = -0.5
    ACCELY_MAX = 0.5
    ACCELX_OPTIONS = 5
    ACCELY_OPTIONS = 5
    CBF_RATE = 3.0
    ENGAGEMENT_DISTANCE = 1.0
    ENGAGEMENT_DISTANCE_REFERENCE_SEPARATION_DISTANCE = 0.5
    DT = 0.1
    DISTANCE_TO_GOAL_THRESHOLD = 0.3
    GOAL_HEADING_THRESHOLD = math.pi / 4
    GOAL_SPEED_THRESHOLD = 0.15
    SEPARATION_DISTANCE = 0.5
    COORDINATION_RANGE = 4
    VALUE_FUNCTION_FILE_NAME = 'data/crazyflies_value_function.pkl'


class RewardWeightConfig:
    MIN_REWARD = -40
    MAX_REWARD = 50
    GOAL_REACH = 50
    SAFETY_VIOLATION = -20
    HJ_VALUE = -2
    POTENTIAL_CONFLICT = -1
    DIFF_FROM_FILTERED_ACTION = -1


class RewardBinaryConfig:
    SAFETY_VIOLATION = False
    HJ_VALUE = False
    POTENTIAL_CONFLICT = False
    SEPARATION_DISTANCE_CURRICULUM = False
    INITIAL_PHASE_USE_SAFETY_FILTER = False
    DIFF_FROM_FILTERED_ACTION = False


DYN_DOUBLE_INTEGRATOR = 0
DYN_AIRTAXI = 1

# reward / curriculum switch bits (match LSM_FLAG_* in include/lsm_b200.h)
FLAG_SAFETY_VIOLATION = 1 << 0
FLAG_HJ_VALUE = 1 << 1
FLAG_POTENTIAL_CONFLICT = 1 << 2
FLAG_SEPARATION_DISTANCE_CURRICULUM = 1 << 3
FLAG_INITIAL_PHASE_USE_SAFETY_FILTER = 1 << 4
FLAG_DIFF_FROM_FILTERED_ACTION = 1 << 5
FLAG_USE_SAFETY_FILTER = 1 << 6   # the --use_safety_filter ARGUMENT (not the per-episode world flag)
FLAG_SHARED_REWARD = 1 << 7       # --collaborative
FLAG_USE_MASKING = 1 << 8         # --use_masking
FLAG_GRAPH_FEAT_GLOBAL = 1 << 9   # --graph_feat_type global
FLAG_INTERP_FLOAT32 = 1 << 10     # grid interpolation in float32 (jax without x64) instead of float64; args.interp_float32


def reward_flags_from(binary_cfg=RewardBinaryConfig) -> int:
    """Capture the class-attribute switches of a RewardBinaryConfig-like object."""
    bits = 0
    if binary_cfg.SAFETY_VIOLATION:
        bits |= FLAG_SAFETY_VIOLATION
    if binary_cfg.HJ_VALUE:
        bits |= FLAG_HJ_VALUE
    if binary_cfg.POTENTIAL_CONFLICT:
        bits |= FLAG_POTENTIAL_CONFLICT
    if binary_cfg.SEPARATION_DISTANCE_CURRICULUM:
        bits |= FLAG_SEPARATION_DISTANCE_CURRICULUM
    if binary_cfg.INITIAL_PHASE_USE_SAFETY_FILTER:
        bits |= FLAG_INITIAL_PHASE_USE_SAFETY_FILTER
    if binary_cfg.DIFF_FROM_FILTERED_ACTION:
        bits |= FLAG_DIFF_FROM_FILTERED_ACTION
    return bits


@dataclass(frozen=True)
class ScenarioParams:
    """Everything `SafeAamScenario.make_world` derives from `args` + config classes
    (reference `navigation_graph_safe.py:100-211`), flattened for the C-ABI."""
    dynamics: int
    num_agents: int
    num_landmarks: int            # landmarks PER AGENT (L)
    episode_length: int
    num_total_episode: int
    num_internal_step: int
    world_size: float
    flags: int
    # dynamics constants
    dt: float
    coordination_range: float     # == max_edge_dist
    dist_thresh: float            # DISTANCE_TO_GOAL_THRESHOLD
    heading_thresh: float         # 0.5 - 0.5 cos(GOAL_HEADING_THRESHOLD)
    speed_thresh: float           # GOAL_SPEED_THRESHOLD
    goal_speed_min: float
    goal_speed_max: float
    separation_distance_target: float
    engagement_distance_ref: float
    engagement_ref_separation: float
    cbf_rate: float
    agent_max_speed: float        # agent.max_speed (min_time)
    # reward weights
    goal_rew: float
    safety_violation_rew: float
    hj_value_rew: float
    potential_conflict_rew: float
    diff_from_filtered_action_rew: float
    min_reward: float
    max_reward: float
    # DECLARED EXTENSION (BASELINE config 3 '+ obstacles'): the reference raises for num_obstacles > 0, see
    # scenario_params_from_args. Entities are agents, landmarks, obstacles (core.py:489-496).
    num_obstacles: int = 0

    @property
    def num_entities(self) -> int:
        return self.num_agents * (1 + self.num_landmarks) + self.num_obstacles

    @property
    def obs_dim(self) -> int:
        return 7 if self.dynamics == DYN_DOUBLE_INTEGRATOR else 6

    @property
    def node_feat_dim(self) -> int:
        if self.flags & FLAG_GRAPH_FEAT_GLOBAL:     # [vel, pos, goal_pos, type], navigation_graph_safe.py:1017-1036
            return 7
        return 10 if self.dynamics == DYN_DOUBLE_INTEGRATOR else 11

    def asdict(self):
        return asdict(self)


# Dynamics limits the kernels carry as COMPILED-IN constants (csrc/lsm_step_common.cuh integrate<> / filter_resolve<>,
# vec_env.action_tables): the reference reads them from its editable config classes (multiagent/config.py:3-60,
# core.py:84-160, safety_filter.py:205-215,380-392). Editing the mirrors above without rebuilding the kernels would give a
# silently inconsistent simulator, so construction checks them.
_COMPILED_LIMITS = {
    'AirTaxiConfig': dict(V_MIN=60 * 0.514444 * 0.001, V_MAX=175 * 0.514444 * 0.001, ACCEL_MIN=-0.001, ACCEL_MAX=0.002,
                          ANGULAR_RATE_MAX=0.1, MOTION_PRIM_ACCEL_OPTIONS=5, MOTION_PRIM_ANGRATE_OPTIONS=5),
    'DoubleIntegratorConfig': dict(VX_MIN=-0.5, VX_MAX=0.5, VY_MIN=-0.5, VY_MAX=0.5, ACCELX_MIN=-0.5, ACCELX_MAX=0.5,
                                   ACCELY_MIN=-0.5, ACCELY_MAX=0.5, ACCELX_OPTIONS=5, ACCELY_OPTIONS=5),
}


def assert_compiled_limits():
    """Raise if a dynamics limit of the config classes differs from the constant compiled into the CUDA kernels."""
    for cls in (AirTaxiConfig, DoubleIntegratorConfig):
        for name, want in _COMPILED_LIMITS[cls.__name__].items():
            have = getattr(cls, name)
            if have != want:
                raise ValueError(f"{cls.__name__}.{name} = {have!r} differs from the value compiled into the kernels ({want!r}): "
                                 f"the dynamics limits are compile-time constants of csrc/lsm_step_common.cuh - change them "
                                 f"there and rebuild")


def scenario_params_from_args(args, binary_cfg=RewardBinaryConfig,
                              weight_cfg=RewardWeightConfig) -> ScenarioParams:
    """Build ScenarioParams from the same argparse Namespace `make_world` consumes."""
    assert_compiled_limits()
    dyn_name = args.dynamics_type
    if dyn_name == 'double_integrator':
        dyn, cfg = DYN_DOUBLE_INTEGRATOR, DoubleIntegratorConfig
        agent_max_speed = DoubleIntegratorConfig.VX_MAX   # core.py:157
    elif dyn_name == 'airtaxi':
        dyn, cfg = DYN_AIRTAXI, AirTaxiConfig
        agent_max_speed = AirTaxiConfig.V_MAX             # core.py:323
    else:
        raise NotImplementedError(f"dynamics_type {dyn_name!r}")
    num_obstacles = int(getattr(args, 'num_obstacles', 0))
    if num_obstacles != 0 and not bool(getattr(args, 'obstacle_extension', False)):
        # the reference itself raises for obstacles in this scenario: 'relative' node features at
        # navigation_graph_safe.py:1064-1065,1087, 'global' one statement later on the N(1+L)-entry disconnect mask (:975-989)
        raise ValueError("obstacle 0 not supported")
    # args.obstacle_extension=True (not a reference argument) opts into the DECLARED extension of SURVEY.md 8c: obstacles are
    # placed, collide (info 'Num_obst_collisions') and enter the distance matrix exactly as the reference's own code does
    # (:230-236, 402-404, 452-465, 1204-1249, core.py:489-543); the two statements that raise are completed as: an obstacle is
    # never disconnected, and its 'relative' node features are the landmark builders' with heading 0, speed 0, entity type 2.
    if num_obstacles < 0 or num_obstacles > 32:
        raise ValueError("num_obstacles must be in [0, 32]")
    if int(getattr(args, 'num_scripted_agents', 0)) != 0 or int(getattr(args, 'num_walls', 0)) != 0:
        raise NotImplementedError("scripted agents / walls are not part of the shipped scenario")
    graph_feat_type = getattr(args, 'graph_feat_type', 'relative')
    if graph_feat_type not in ('relative', 'global'):
        raise ValueError(f"graph_feat_type {graph_feat_type!r}")
    if not bool(getattr(args, 'use_masking', True)):
        # reference quirk Q9: without masking a parked agent overruns its landmark list and get_entity raises
        raise NotImplementedError("use_masking=False makes the reference raise once an agent parks (Q9)")
    num_landmarks = int(args.num_landmarks)
    if num_landmarks < 2:
        # creat_relative_heading_list_from_goal_position_list asserts len > 1 (utils.py:31)
        raise AssertionError("Goal position list should have more than 1 element")
    flags = reward_flags_from(binary_cfg) | FLAG_USE_MASKING
    if bool(args.use_safety_filter):
        flags |= FLAG_USE_SAFETY_FILTER
    if bool(getattr(args, 'collaborative', False)):
        flags |= FLAG_SHARED_REWARD
    if graph_feat_type == 'global':
        flags |= FLAG_GRAPH_FEAT_GLOBAL
    # not a reference argument: which DECLARED arithmetic of hj_reachability's Grid.interpolate to reproduce
    if bool(getattr(args, 'interp_float32', False)):
        flags |= FLAG_INTERP_FLOAT32
    num_total_episode = int(args.num_env_steps) // int(args.episode_length) // int(args.n_rollout_threads)
    return ScenarioParams(
        dynamics=dyn,
        num_agents=int(args.num_agents),
        num_landmarks=num_landmarks,
        episode_length=int(args.episode_length),
        num_total_episode=num_total_episode,
        num_internal_step=int(getattr(args, 'num_internal_step', 1)),
        world_size=float(args.world_size),
        flags=flags,
        dt=float(cfg.DT),
        coordination_range=float(cfg.COORDINATION_RANGE),
        dist_thresh=float(cfg.DISTANCE_TO_GOAL_THRESHOLD),
        heading_thresh=0.5 - 0.5 * math.cos(cfg.GOAL_HEADING_THRESHOLD),
        speed_thresh=float(cfg.GOAL_SPEED_THRESHOLD),
        goal_speed_min=float(cfg.V_MIN),
        goal_speed_max=float(cfg.V_NOMINAL),
        separation_distance_target=float(cfg.SEPARATION_DISTANCE),
        engagement_distance_ref=float(cfg.ENGAGEMENT_DISTANCE),
        engagement_ref_separation=float(cfg.ENGAGEMENT_DISTANCE_REFERENCE_SEPARATION_DISTANCE),
        cbf_rate=float(cfg.CBF_RATE),
        agent_max_speed=float(agent_max_speed),
        goal_rew=float(weight_cfg.GOAL_REACH),
        safety_violation_rew=float(weight_cfg.SAFETY_VIOLATION),
        hj_value_rew=float(weight_cfg.HJ_VALUE),
        potential_conflict_rew=float(weight_cfg.POTENTIAL_CONFLICT),
        diff_from_filtered_action_rew=float(weight_cfg.DIFF_FROM_FILTERED_ACTION),
        min_reward=float(weight_cfg.MIN_REWARD),
        max_reward=float(weight_cfg.MAX_REWARD),
        num_obstacles=num_obstacles,
    )
