"""HBM data layout of the batched simulator (must match include/lsm_b200.h).

State is struct-of-arrays over (field, env, agent): a warp that owns a group of consecutive
environments reads each field as one contiguous, coalesced run of doubles.

    agent_f64 : (AF_COUNT, num_envs, N)      float64
    agent_i32 : (AI_COUNT, num_envs, N)      int32
    landmarks : (LF_COUNT, num_envs, N*L)    float64   landmark m = order * N + agent
    obstacles : (2, num_envs, O)             float64   obstacle x / y (obstacle extension; O = 0 in every shipped script)
    env_f64   : (EF_COUNT, num_envs)         float64
    env_i32   : (EI_COUNT, num_envs)         int32

`times_required` and `dists_to_goal` (world arrays of the reference's info_callback,
navigation_graph_safe.py:392-401) are ping-ponged between an A and a B slot: slot
`EI_PARITY` holds the newest values, the other slot the previous step's. The lazily built
`infos` need both, because agent i's info dict sees the new value of agents <= i and the
old value of agents > i (sequential loop, environment.py:979-1029).
"""
from __future__ import annotations

# per-agent float64 fields
(AF_X, AF_Y, AF_S2, AF_S3, AF_P_DIST, AF_STATE_TIME, AF_MIN_REL_DIST, AF_GOAL_MIN_TIME,
 AF_TIMES_REQ_A, AF_TIMES_REQ_B, AF_DISTS_GOAL_A, AF_DISTS_GOAL_B, AF_DIST_LEFT,
 AF_EP_TRAVEL_DIST, AF_EP_MIN_DIST, AF_ACTION_DIFF) = range(16)
AF_COUNT = 16

# per-agent int32 fields
(AI_REACHED, AI_DONE, AI_SAFETY_FILTERED, AI_DECONFLICT_IDX, AI_NUM_COLLISIONS,
 AI_EP_TRAVEL_LEN, AI_EP_CONFLICT, AI_EP_MULTI, AI_EP_DONE, AI_NUM_OBST_COLLISIONS) = range(10)
AI_COUNT = 10

# per-landmark float64 fields
LF_X, LF_Y, LF_HEADING, LF_SPEED, LF_SIN, LF_COS = range(6)
LF_COUNT = 6

EF_CURRICULUM_RATIO = 0
EF_COUNT = 1

EI_CURRENT_STEP, EI_RESET_COUNT, EI_PARITY, EI_JUST_RESET = range(4)
EI_COUNT = 4

# terminal-step info snapshot of auto-resetting envs (include/lsm_b200.h LSM_TF_* / LSM_TI_*)
(TF_X, TF_Y, TF_MIN_REL_DIST, TF_DIST_LEFT, TF_TIMES_REQ_NEW, TF_TIMES_REQ_OLD, TF_DISTS_GOAL_NEW, TF_DISTS_GOAL_OLD,
 TF_GOAL_MIN_TIME) = range(9)
TF_COUNT = 9
TI_NUM_COLLISIONS, TI_SAFETY_FILTERED, TI_NUM_OBST_COLLISIONS = range(3)
TI_COUNT = 3

# episode summary (environment.py:1065-1073), in this order
EP_INFO_KEYS = ('travel_time_mean', 'travel_distance_mean', 'done_percentage', 'num_reached_goal_mean',
                'conflict_percentage', 'min_distance_mean', 'min_distance_min',
                'multiple_engagement_percentage')
EP_COUNT = 8

NUM_ACTIONS = 25
MAX_AGENTS = 32
MAX_LANDMARKS = 128   # np.int8 landmark index in the reference (Q8)
MAX_OBSTACLES = 32
