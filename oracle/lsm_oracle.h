/*
 * lsm_oracle.h - CPU restatement (plain C, float64, no FMA contraction) of the per-step hot
 * path of Layered-Safe-MARL's `navigation_graph_safe` environment.
 *
 * TEST INFRASTRUCTURE. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this library. The product (layered_safe_marl_b200/) never
 * links, imports or calls it.
 *
 * Pinning: checked against golden rollouts produced by running the UNMODIFIED reference
 * (/root/reference/multiagent) through oracle/gen_golden.py (fixtures under tests/golden/).
 * Third-party arithmetic that is absent offline (hj_reachability 0.5.0 grid interpolation,
 * hj_reachability_utils, cvxpy 1.4.1/OSQP) follows the declared semantics of
 * oracle/ref_stubs - for those pieces parity with the real libraries is UNPINNED.
 *
 * Memory layout is struct-of-arrays over (field, env, agent); see the LSMO_* indices.
 */
#ifndef LSM_ORACLE_H
#define LSM_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { LSMO_DYN_DI = 0, LSMO_DYN_AIRTAXI = 1 };

enum {
    LSMO_FLAG_SAFETY_VIOLATION = 1 << 0,
    LSMO_FLAG_HJ_VALUE = 1 << 1,
    LSMO_FLAG_POTENTIAL_CONFLICT = 1 << 2,
    LSMO_FLAG_SEPARATION_DISTANCE_CURRICULUM = 1 << 3,
    LSMO_FLAG_INITIAL_PHASE_USE_SAFETY_FILTER = 1 << 4,
    LSMO_FLAG_DIFF_FROM_FILTERED_ACTION = 1 << 5,
    LSMO_FLAG_USE_SAFETY_FILTER = 1 << 6,
    LSMO_FLAG_SHARED_REWARD = 1 << 7,
    LSMO_FLAG_USE_MASKING = 1 << 8,
    LSMO_FLAG_GRAPH_FEAT_GLOBAL = 1 << 9,   /* --graph_feat_type global (navigation_graph_safe.py:1017-1036) */
    LSMO_FLAG_INTERP_FLOAT32 = 1 << 10      /* Grid.interpolate in float32 (jax default dtype) instead of float64 */
};

/* per-agent float64 fields: agent_f64[field][env][agent] */
enum {
    LSMO_AF_X = 0, LSMO_AF_Y, LSMO_AF_S2 /* vx | theta */, LSMO_AF_S3 /* vy | speed */,
    LSMO_AF_P_DIST, LSMO_AF_STATE_TIME, LSMO_AF_MIN_REL_DIST, LSMO_AF_GOAL_MIN_TIME,
    LSMO_AF_TIMES_REQ_A, LSMO_AF_TIMES_REQ_B, LSMO_AF_DISTS_GOAL_A, LSMO_AF_DISTS_GOAL_B,
    LSMO_AF_DIST_LEFT, LSMO_AF_EP_TRAVEL_DIST, LSMO_AF_EP_MIN_DIST, LSMO_AF_ACTION_DIFF,
    LSMO_AF_COUNT
};
/* per-agent int32 fields: agent_i32[field][env][agent] */
enum {
    LSMO_AI_REACHED = 0, LSMO_AI_DONE, LSMO_AI_SAFETY_FILTERED, LSMO_AI_DECONFLICT_IDX,
    LSMO_AI_NUM_COLLISIONS, LSMO_AI_EP_TRAVEL_LEN, LSMO_AI_EP_CONFLICT, LSMO_AI_EP_MULTI,
    LSMO_AI_EP_DONE, LSMO_AI_NUM_OBST_COLLISIONS /* world.num_obstacle_collisions (obstacle extension) */, LSMO_AI_COUNT
};
/* per-landmark float64 fields: landmarks[field][env][l*N + agent] */
enum { LSMO_LF_X = 0, LSMO_LF_Y, LSMO_LF_HEADING, LSMO_LF_SPEED, LSMO_LF_SIN, LSMO_LF_COS, LSMO_LF_COUNT };
/* per-env fields */
enum { LSMO_EF_CURRICULUM_RATIO = 0, LSMO_EF_COUNT };
enum { LSMO_EI_CURRENT_STEP = 0, LSMO_EI_RESET_COUNT, LSMO_EI_PARITY, LSMO_EI_JUST_RESET, LSMO_EI_COUNT };
/* episode summary written at reset: ep_info[env][k] (environment.py:1065-1073) */
enum {
    LSMO_EP_TRAVEL_TIME_MEAN = 0, LSMO_EP_TRAVEL_DISTANCE_MEAN, LSMO_EP_DONE_PERCENTAGE,
    LSMO_EP_NUM_REACHED_GOAL_MEAN, LSMO_EP_CONFLICT_PERCENTAGE, LSMO_EP_MIN_DISTANCE_MEAN,
    LSMO_EP_MIN_DISTANCE_MIN, LSMO_EP_MULTIPLE_ENGAGEMENT_PERCENTAGE, LSMO_EP_COUNT
};

typedef struct lsmo_params {
    int32_t dynamics, num_agents, num_landmarks, episode_length;
    int32_t num_total_episode, num_internal_step;
    uint32_t flags;
    int32_t num_obstacles;   /* --num_obstacles. > 0 is the DECLARED EXTENSION of SURVEY 8c (the reference raises): entities are
                                agents, landmarks, obstacles (core.py:489-496); obstacles are never disconnected; an obstacle's
                                node features are the landmark builders' with heading 0, speed 0, entity type 2 */
    double world_size;
    double dt, coordination_range, dist_thresh, heading_thresh, speed_thresh;
    double goal_speed_min, goal_speed_max, separation_distance_target;
    double engagement_distance_ref, engagement_ref_separation, cbf_rate, agent_max_speed;
    double goal_rew, safety_violation_rew, hj_value_rew, potential_conflict_rew;
    double diff_from_filtered_action_rew, min_reward, max_reward;
    double act_tab0[5], act_tab1[5];    /* np.linspace tables of environment.py:387-410 */
} lsmo_params;

typedef struct lsmo_grid {
    int32_t ndim;
    int32_t shape[5];
    int32_t periodic[5];
    int32_t _pad;
    double lo[5], hi[5];
    double separation_distance;   /* separation the values encode (HjDataHandle.separation_distance) */
    double ttr_max;
    const float *values;          /* [prod(shape)] C order */
    const float *grads;           /* [prod(shape)][ndim] or NULL */
} lsmo_grid;

typedef struct lsmo_buffers {
    int64_t num_envs;
    int64_t env_id_base;   /* global index of env 0 of this shard (keys the reset RNG stream) */
    double *agent_f64;     /* [LSMO_AF_COUNT][num_envs][N] */
    int32_t *agent_i32;    /* [LSMO_AI_COUNT][num_envs][N] */
    double *landmarks;     /* [LSMO_LF_COUNT][num_envs][N*L] */
    double *env_f64;       /* [LSMO_EF_COUNT][num_envs] */
    int32_t *env_i32;      /* [LSMO_EI_COUNT][num_envs] */
    /* outputs */
    float *obs;            /* [num_envs][N][D] */
    float *node_obs;       /* [num_envs][N][E][F]   E = N(1+L) + O */
    float *adj;            /* [num_envs][N][E][E] */
    float *reward;         /* [num_envs][N] */
    uint8_t *done;         /* [num_envs][N] */
    double *safe_action;   /* [num_envs][N][2] applied (filtered) control of the last internal step */
    double *ep_info;       /* [num_envs][LSMO_EP_COUNT] */
    double *obstacles;     /* [2][num_envs][O] obstacle x / y (NULL when num_obstacles == 0) */
} lsmo_buffers;

/* One env.step (+ graphworker auto-reset when `auto_reset`) for every env.
 * action_idx: [num_envs][N] int32 in [0,25). episode: the `num_current_episode` passed to step.
 * Returns 0 on success. nthreads<=1 -> serial. */
int lsmo_step(const lsmo_params *p, const lsmo_grid *value_grid, const lsmo_grid *ttr_grid,
              const lsmo_buffers *b, const int32_t *action_idx, int64_t episode, uint64_t seed,
              int auto_reset, int nthreads);

/* env.reset(episode) for every env with env_mask[e] != 0 (NULL -> all). sample!=0 draws a new
 * random scenario (Philox stream keyed by seed/env/reset_count); sample==0 keeps the injected
 * agent/landmark state and only runs the bookkeeping + observation emission of reset. */
int lsmo_reset(const lsmo_params *p, const lsmo_grid *value_grid, const lsmo_grid *ttr_grid,
               const lsmo_buffers *b, const uint8_t *env_mask, int64_t episode, uint64_t seed,
               int sample, int nthreads);

/* Re-emit obs/node_obs/adj from the current state without stepping (used after set_state). */
int lsmo_observe(const lsmo_params *p, const lsmo_buffers *b, int nthreads);

/* helpers exported for unit tests */
double lsmo_interpolate(const lsmo_grid *g, const double *x, int component /* -1: values */);
void lsmo_set_interp_float32(int on);   /* arithmetic of lsmo_interpolate when called directly (the entry points set it from params.flags) */
void lsmo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void lsmo_curriculum(const lsmo_params *p, double ratio, double out[12]);
double lsmo_magnetic_heading(double px, double py, double radius);
/* include/lsm_math.h evaluated on the host (op: 0 sin, 1 cos, 2 atan2(a, b)) */
void lsmo_math_eval(int op, const double *a, const double *b, double *out, int64_t n);
/* airtaxi relative position (safety_filter.py:277-284): 0 = literal atan2 / cos / sin form (default), 1 = rotation
 * form (what the specialised CUDA pipeline evaluates; differs from the literal form by ~1e-16 relative). Process-global. */
void lsmo_set_relative_state_form(int form);
int lsmo_get_relative_state_form(void);

#ifdef __cplusplus
}
#endif
#endif
