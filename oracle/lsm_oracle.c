/*
 * lsm_oracle.c - sequential CPU restatement of the navigation_graph_safe step path.
 * See lsm_oracle.h for scope and pinning. Every function cites the reference file:line it
 * follows (paths relative to /root/reference). Build: oracle/Makefile (gcc -O2
 * -ffp-contract=off, so a*b+c is never fused: every float64 intermediate rounds exactly as the
 * reference's numpy scalar arithmetic does).
 *
 * The control flow is deliberately the reference's own: a sequential per-agent loop that
 * mutates goal counters / done flags / velocities in place while observations are being
 * emitted (multiagent/environment.py:979-1029). The CUDA kernels use a parallel pre/post
 * formulation instead; agreement between the two is part of what the parity tests check.
 */
#include "lsm_oracle.h"
/* float64 sin / cos / atan2 shared bit for bit with the CUDA kernels (numpy's libm in the reference) */
#include "../include/lsm_math.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define MAXN 32
#define MAXM 128
#define MAXO 32
#define MAXE (MAXN + MAXM + MAXO)
#define PI 3.141592653589793

static inline double clipd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }
/* python builtins max(a,b)/min(a,b): return the first argument unless the second compares greater/less */
static inline double pymax(double a, double b) { return (b > a) ? b : a; }
static inline double pymin(double a, double b) { return (b < a) ? b : a; }
static inline double norm2(double a, double b) { return sqrt(a * a + b * b); }

/* ------------------------------------------------------------------------------------------
 * Philox4x32-10 (Salmon et al. 2011) - counter based, so oracle and CUDA draw identical numbers.
 * ---------------------------------------------------------------------------------------- */
void lsmo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

typedef struct {
    uint32_t key[2];
    uint32_t env, reset_count;
    uint32_t block;     /* next block index */
    uint32_t buf[4];
    int have;           /* doubles left in buf (0..2) */
} rng_t;

static void rng_init(rng_t *r, uint64_t seed, uint32_t env, uint32_t reset_count) {
    r->key[0] = (uint32_t)seed; r->key[1] = (uint32_t)(seed >> 32);
    r->env = env; r->reset_count = reset_count; r->block = 0; r->have = 0;
}
/* 53-bit uniform in [0,1): (a>>5, b>>6) like numpy's legacy random_double */
static double rng_uniform01(rng_t *r) {
    if (r->have == 0) {
        uint32_t ctr[4] = { r->block, r->env, r->reset_count, 0u };
        lsmo_philox4x32_10(ctr, r->key, r->buf);
        r->block++; r->have = 2;
    }
    int k = 2 - r->have;
    r->have--;
    uint32_t a = r->buf[2 * k] >> 5, b = r->buf[2 * k + 1] >> 6;
    return ((double)a * 67108864.0 + (double)b) / 9007199254740992.0;
}
/* np.random.uniform(lo, hi) = lo + (hi - lo) * U */
static double rng_uniform(rng_t *r, double lo, double hi) { return lo + (hi - lo) * rng_uniform01(r); }

/* ------------------------------------------------------------------------------------------
 * Grid interpolation - declared semantics of oracle/ref_stubs/hj_reachability (Grid.interpolate);
 * call sites multiagent/safety_filter.py:195,245,348,418, core.py:463, navigation_graph_safe.py:751.
 * ---------------------------------------------------------------------------------------- */
/* LSMO_FLAG_INTERP_FLOAT32 (set from the params at every entry point): position, floor, weights, weight products and the
 * corner sum in float32 - what jax computes without jax_enable_x64, which the reference never enables - instead of the
 * same formula in float64. Same corner order, same index clamping / wrapping. */
static int g_interp_f32 = 0;
void lsmo_set_interp_float32(int on) { g_interp_f32 = on ? 1 : 0; }      /* unit tests of lsmo_interpolate */

static double interpolate_f32(const lsmo_grid *g, const double *x, int component) {
    int nd = g->ndim;
    int64_t idx[2][5];
    float w[2][5];
    for (int d = 0; d < nd; ++d) {
        double n = (double)g->shape[d];
        double spacing = g->periodic[d] ? (g->hi[d] - g->lo[d]) / n : (g->hi[d] - g->lo[d]) / (n - 1.0);
        float pos = ((float)x[d] - (float)g->lo[d]) / (float)spacing;
        if (isnan(pos)) return NAN;
        pos = fminf(fmaxf(pos, -1.0e9f), 1.0e9f);
        float fl = floorf(pos);
        float whi = pos - fl;
        w[0][d] = 1.0f - whi; w[1][d] = whi;
        int64_t il = (int64_t)fl, ih = il + 1, s = g->shape[d];
        if (g->periodic[d]) {
            il %= s; if (il < 0) il += s;
            ih %= s; if (ih < 0) ih += s;
        } else {
            il = il < 0 ? 0 : (il > s - 1 ? s - 1 : il);
            ih = ih < 0 ? 0 : (ih > s - 1 ? s - 1 : ih);
        }
        idx[0][d] = il; idx[1][d] = ih;
    }
    float acc = 0.0f;
    for (int corner = 0; corner < (1 << nd); ++corner) {
        float weight = 0.0f; int64_t lin = 0;
        for (int d = 0; d < nd; ++d) {
            int bit = (corner >> (nd - 1 - d)) & 1;
            weight = (d == 0) ? w[bit][d] : weight * w[bit][d];
            lin = lin * g->shape[d] + idx[bit][d];
        }
        float v = (component < 0) ? g->values[lin] : g->grads[lin * nd + component];
        acc = acc + weight * v;
    }
    return (double)acc;
}

double lsmo_interpolate(const lsmo_grid *g, const double *x, int component) {
    if (g_interp_f32) return interpolate_f32(g, x, component);
    int nd = g->ndim;
    int64_t idx[2][5];
    double w[2][5];
    for (int d = 0; d < nd; ++d) {
        double n = (double)g->shape[d];
        double spacing = g->periodic[d] ? (g->hi[d] - g->lo[d]) / n : (g->hi[d] - g->lo[d]) / (n - 1.0);
        double pos = (x[d] - g->lo[d]) / spacing;
        if (isnan(pos)) return NAN;
        pos = clipd(pos, -1.0e9, 1.0e9);
        double fl = floor(pos);
        double whi = pos - fl;
        w[0][d] = 1.0 - whi; w[1][d] = whi;
        int64_t il = (int64_t)fl, ih = il + 1, s = g->shape[d];
        if (g->periodic[d]) {
            il %= s; if (il < 0) il += s;
            ih %= s; if (ih < 0) ih += s;
        } else {
            il = il < 0 ? 0 : (il > s - 1 ? s - 1 : il);
            ih = ih < 0 ? 0 : (ih > s - 1 ? s - 1 : ih);
        }
        idx[0][d] = il; idx[1][d] = ih;
    }
    double acc = 0.0;
    for (int corner = 0; corner < (1 << nd); ++corner) {
        double weight = 0.0; int64_t lin = 0;
        for (int d = 0; d < nd; ++d) {
            int bit = (corner >> (nd - 1 - d)) & 1;
            weight = (d == 0) ? w[bit][d] : weight * w[bit][d];
            lin = lin * g->shape[d] + idx[bit][d];
        }
        double v = (component < 0) ? (double)g->values[lin] : (double)g->grads[lin * nd + component];
        acc = acc + weight * v;
    }
    return acc;
}

/* ------------------------------------------------------------------------------------------
 * Curriculum scalars - navigation_graph_safe.py:324-366 (update_curriculum),
 * :1101-1122 (effective ratios), :319-322 (engagement distance).
 * out: 0 sloped, 1 stair, 2 heading_thresh, 3 speed_thresh, 4 dist_thresh, 5 multi_eng_rew_scaled,
 *      6 conflict_rew_scaled, 7 diff_rew_scaled, 8 conflict_value_rew_scaled,
 *      9 separation_distance, 10 engagement_distance, 11 world.use_safety_filter (0/1)
 * ---------------------------------------------------------------------------------------- */
static double ratio_sloped(double ratio, double start, double end) {
    return clipd(ratio - start, 0.0, end - start) / (end - start);
}
static double ratio_stair(double ratio, int num_steps, double start, double end) {
    if (ratio < start) return 0.0;
    if (ratio > end) return 1.0;
    double cont = (double)(num_steps - 1) * clipd(ratio - start, 0.0, end - start) / (end - start);
    return (1.0 + floor(cont)) / (double)num_steps;
}
void lsmo_curriculum(const lsmo_params *p, double ratio, double out[12]) {
    double sloped = ratio_sloped(ratio, 0.25, 0.75);
    double stair = ratio_stair(ratio, 4, 0.2, 0.75);
    out[0] = sloped; out[1] = stair;
    out[2] = p->heading_thresh * (1.0 - sloped) + p->heading_thresh * sloped;
    out[3] = p->speed_thresh * (1.0 - stair) + p->speed_thresh * stair;
    out[4] = p->dist_thresh * (1.0 - stair) + p->dist_thresh * stair;
    out[5] = p->potential_conflict_rew * stair;
    out[6] = p->safety_violation_rew * stair;
    out[7] = p->diff_from_filtered_action_rew * stair;
    out[8] = p->hj_value_rew * stair;
    double phase = ratio_stair(ratio, 4, 0.2, 0.75) * 0.5 * PI;
    double sep_ratio = 1.0 - lsm_cos(phase);
    int use_filter_arg = (p->flags & LSMO_FLAG_USE_SAFETY_FILTER) != 0;
    int initial_phase = use_filter_arg && (p->flags & LSMO_FLAG_INITIAL_PHASE_USE_SAFETY_FILTER);
    int world_filter = use_filter_arg;
    if (!initial_phase && use_filter_arg) world_filter = sloped > 0.0;
    double sep_init = (p->flags & LSMO_FLAG_SEPARATION_DISTANCE_CURRICULUM) ? 0.0 : p->separation_distance_target;
    double sep = sep_init * (1.0 - sep_ratio) + p->separation_distance_target * sep_ratio;
    out[9] = sep;
    out[10] = p->engagement_distance_ref + (sep - p->engagement_ref_separation);
    out[11] = (double)world_filter;
}

/* ------------------------------------------------------------------------------------------
 * Magnetic-field reference heading - multiagent/custom_scenarios/utils.py:276-321
 * ---------------------------------------------------------------------------------------- */
double lsmo_magnetic_heading(double px, double py, double radius) {
    if (fabs(px) < 1e-6) return 0.0;
    const double scale_x = 0.5;
    px = scale_x * px;
    const int NSEG = 50;
    const double step = (2.0 * PI - 0.0) / (double)NSEG;   /* np.linspace(0, 2pi, 50, endpoint=False) */
    double bx = 0.0, by = 0.0;
    for (int k = 0; k < NSEG; ++k) {
        double phi = (double)k * step + 0.0;
        double c = lsm_cos(phi), s = lsm_sin(phi);
        double Ly = -radius * c, Lz = -radius * s;
        double dLy = radius * s, dLz = -radius * c;
        double rx = px - 0.0, ry = py - Ly, rz = 0.0 - Lz;
        double rmag = sqrt((rx * rx + ry * ry) + rz * rz);
        double rmag3 = rmag * rmag * rmag;
        double cx = dLy * rz - dLz * ry;
        double cy = dLz * rx - 0.0 * rz;
        bx = bx + cx / rmag3;
        by = by + cy / rmag3;
    }
    bx = bx / scale_x;
    return lsm_atan2(by, bx);
}

/* ------------------------------------------------------------------------------------------
 * Per-env working set
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    const lsmo_params *p;
    const lsmo_grid *vg, *tg;
    int N, L, M, O, E, dyn;
    double cur[12];          /* curriculum scalars */
    double ratio;
    /* agents */
    double x[MAXN], y[MAXN], s2[MAXN], s3[MAXN];
    double p_dist[MAXN], state_time[MAXN], min_rel[MAXN], goal_min_time[MAXN];
    double times_req_old[MAXN], times_req[MAXN], dists_goal_old[MAXN], dists_goal[MAXN], dist_left[MAXN];
    double ep_travel_dist[MAXN], ep_min_dist[MAXN], action_diff[MAXN];
    int reached[MAXN], done[MAXN], safety_filtered[MAXN], deconflict[MAXN], ncoll[MAXN];
    int ep_travel_len[MAXN], ep_conflict[MAXN], ep_multi[MAXN], ep_done[MAXN];
    /* landmarks */
    double lx[MAXM], ly[MAXM], lh[MAXM], ls[MAXM], lsin[MAXM], lcos[MAXM];
    /* obstacles (declared extension, lsm_oracle.h: num_obstacles) */
    double ox[MAXO], oy[MAXO];
    int nobst[MAXN];         /* world.num_obstacle_collisions */
    int current_step, reset_count, parity;
    double *D;               /* E*E cached_dist_mag (masked in place like the reference) */
} env_t;

static inline double a_theta(const env_t *w, int i) {   /* core.py:97-99,179-181 */
    return w->dyn == LSMO_DYN_DI ? lsm_atan2(w->s3[i], w->s2[i]) : w->s2[i];
}
static inline double a_speed(const env_t *w, int i) {   /* core.py:89-91,174-176 */
    return w->dyn == LSMO_DYN_DI ? sqrt(w->s2[i] * w->s2[i] + w->s3[i] * w->s3[i]) : w->s3[i];
}
static inline void a_vel(const env_t *w, int i, double v[2]) {   /* core.py:105-108,183-185 */
    if (w->dyn == LSMO_DYN_DI) { v[0] = w->s2[i]; v[1] = w->s3[i]; }
    else { v[0] = w->s3[i] * lsm_cos(w->s2[i]); v[1] = w->s3[i] * lsm_sin(w->s2[i]); }
}

/* navigation_graph_safe.py:576-582 get_agent_current_goal (landmark index) */
static inline int current_goal(const env_t *w, int i) {
    int order = w->reached[i] * w->N + i;
    if (order >= w->M) order = (w->reached[i] - 1) * w->N + i;
    return order;
}

/* utils.py:79-81 */
static inline double direction_alignment_error(double h, double href) { return 0.5 - 0.5 * lsm_cos(h - href); }

/* utils.py:104-112: rot = [[c, s], [-s, c]] applied to (q - ref) */
static inline void rel_pos_from_reference(double qx, double qy, double rx, double ry, double heading, double out[2]) {
    double dx = qx - rx, dy = qy - ry;
    double c = lsm_cos(heading), s = lsm_sin(heading);
    out[0] = c * dx + s * dy;
    out[1] = (-s) * dx + c * dy;
}

/* navigation_graph_safe.py:606-656 evaluate_agent_goal_reached (+ DI heading condition) */
static int goal_reached(const env_t *w, int i) {
    int g = current_goal(w, i);
    double dx = w->x[i] - w->lx[g], dy = w->y[i] - w->ly[g];
    double dist = sqrt(dx * dx + dy * dy);
    double th = a_theta(w, i);
    double he = direction_alignment_error(th, w->lh[g]);
    double ve = fabs(a_speed(w, i) - w->ls[g]);
    double heading_thresh = w->cur[2], speed_thresh = w->cur[3], dist_thresh = w->cur[4];
    int cond;
    if (w->dyn == LSMO_DYN_DI) {
        const double speed_advantage_thresh = 0.2;
        if (dist > dist_thresh) cond = he < heading_thresh;
        else if (w->ls[g] > speed_advantage_thresh) cond = he < heading_thresh;
        else {
            double sa = clipd(1.0 - w->ls[g] / speed_advantage_thresh, 0.0, 1.0);
            double tc = 0.5 * sa + heading_thresh * (1.0 - sa);
            double da = clipd(1.0 - dist / dist_thresh, 0.0, 1.0);
            double tca = tc * da + heading_thresh * (1.0 - da);
            cond = he < tca;
        }
    } else cond = he < heading_thresh;
    return dist < dist_thresh && cond && ve < speed_thresh;
}

/* navigation_graph_safe.py:658-675 + 1091-1099 */
static void update_reached_goal_and_done(env_t *w, int i) {
    if (goal_reached(w, i)) {
        if (w->p->flags & LSMO_FLAG_USE_MASKING) { if (!w->done[i]) w->reached[i] += 1; }
        else w->reached[i] += 1;
    }
    if (w->reached[i] >= w->L) {
        w->done[i] = 1;
        if (w->dyn == LSMO_DYN_DI) { w->s2[i] = 0.0; w->s3[i] = 0.0; }
        else w->s3[i] = 0.0;
    }
}

/* value of the relative state between two agents, +inf when NaN
 * (safety_filter.py:192-201,345-354; core.py:459-468) */
static void di_relative_state(const env_t *w, int e, int o, double r[4]) {      /* safety_filter.py:356-362 */
    r[0] = w->x[e] - w->x[o]; r[1] = w->y[e] - w->y[o];
    r[2] = w->s2[e] - w->s2[o]; r[3] = w->s3[e] - w->s3[o];
}
/* LITERAL (default): d cos(phi - theta_e), d sin(phi - theta_e) with phi = atan2(dy, dx), as the reference writes it.
 * ROTATION: the same point as the rotation of (dx, dy) by theta_e,
 *     d cos(phi - theta_e) = dx cos(theta_e) + dy sin(theta_e),   d sin(phi - theta_e) = dy cos(theta_e) - dx sin(theta_e)
 * - equal to the literal form to ~1e-16 d; it is the form the specialised CUDA pipeline evaluates (one sincos per pair
 * instead of sqrt + atan2 + cos + sin). tests/ select it to compare that pipeline BIT FOR BIT; the golden rollouts of
 * the reference are checked with the literal form, and tests/test_oracle_golden.py checks that the two forms give
 * the same discrete outputs on them. */
static int g_rel_form = 0;
void lsmo_set_relative_state_form(int form) { g_rel_form = form ? 1 : 0; }
int lsmo_get_relative_state_form(void) { return g_rel_form; }
static void at_relative_state(const env_t *w, int e, int o, double r[5]) {      /* safety_filter.py:277-284 */
    double ddx = w->x[o] - w->x[e], ddy = w->y[o] - w->y[e];
    double rel_heading = w->s2[o] - w->s2[e];
    if (g_rel_form == 0) {
        double dist = sqrt(ddx * ddx + ddy * ddy);
        double ang = lsm_atan2(ddy, ddx);
        r[0] = dist * lsm_cos(ang - w->s2[e]);
        r[1] = dist * lsm_sin(ang - w->s2[e]);
    } else {
        double se, ce;
        lsm_sincos(w->s2[e], &se, &ce);
        r[0] = ddx * ce + ddy * se;
        r[1] = ddy * ce - ddx * se;
    }
    r[2] = rel_heading; r[3] = w->s3[e]; r[4] = w->s3[o];
}
static double hj_value(const env_t *w, const double *rel, int *in_range) {
    double v = lsmo_interpolate(w->vg, rel, -1);
    if (isnan(v)) { *in_range = 0; return INFINITY; }
    *in_range = 1;
    /* HjDataHandle.update_separation_distance (safety_filter.py:170-174): values shift by the
     * difference between the env's current separation distance and the one the grid encodes */
    return v - (w->cur[9] - w->vg->separation_distance);
}

/* single-constraint QP, declared semantics of oracle/ref_stubs/cvxpy. returns 0 if infeasible
 * (caller then ALIASES u to u_ref like `return u_ref` at safety_filter.py:304-305,373-375). */
static int qp_project(const double a[4], double b, const double r[4], const double pinv[4], double u[4]) {
    double s = 0.0;
    for (int k = 0; k < 4; ++k) s = s + a[k] * r[k];
    s = s + b;
    if (s >= 0.0) { for (int k = 0; k < 4; ++k) u[k] = r[k]; return 1; }
    double denom = 0.0;
    for (int k = 0; k < 4; ++k) denom = denom + (a[k] * pinv[k]) * a[k];
    if (denom == 0.0) return 0;
    double lam = s / denom;
    for (int k = 0; k < 4; ++k) u[k] = r[k] - lam * (pinv[k] * a[k]);
    return 1;
}
static inline double f32r(double v) { return (double)(float)v; }

/* World.apply_safety_filter (core.py:648-677) + the two handles (safety_filter.py:203-260,378-433).
 * raw[i][2] in, safe[i][2] out. */
static void apply_safety_filter(env_t *w, double raw[][2], double safe[][2], int filtered[], int deconf[]) {
    const lsmo_params *p = w->p;
    int N = w->N;
    for (int i = 0; i < N; ++i) {
        safe[i][0] = raw[i][0]; safe[i][1] = raw[i][1]; filtered[i] = 0; deconf[i] = -1;
        if (w->done[i]) continue;
        int others[MAXN], no = 0;
        for (int j = 0; j < N; ++j) if (j != i && !w->done[j]) others[no++] = j;
        if (no == 0) continue;
        double best_d = 0.0, best_v = 0.0; int kd = -1, kv = -1, kv_in_range = 0;
        for (int k = 0; k < no; ++k) {
            int j = others[k];
            double ddx = w->x[j] - w->x[i], ddy = w->y[j] - w->y[i];
            double dist = sqrt(ddx * ddx + ddy * ddy);
            double rel[5]; int inr;
            if (w->dyn == LSMO_DYN_DI) di_relative_state(w, i, j, rel); else at_relative_state(w, i, j, rel);
            double v = hj_value(w, rel, &inr);
            if (kd < 0 || dist < best_d) { kd = k; best_d = dist; }       /* np.argmin: first minimum */
            if (kv < 0 || v < best_v) { kv = k; best_v = v; kv_in_range = inr; }
        }
        int jo = others[kv];
        deconf[i] = jo;
        if (best_d > p->coordination_range) continue;
        if (!kv_in_range) continue;
        double rel[5];
        if (w->dyn == LSMO_DYN_DI) di_relative_state(w, i, jo, rel); else at_relative_state(w, i, jo, rel);
        double uref[4] = { raw[i][0], raw[i][1], raw[jo][0], raw[jo][1] };
        double g[5];
        for (int d = 0; d < w->vg->ndim; ++d) g[d] = lsmo_interpolate(w->vg, rel, d);
        const double eps_hj = 0.4;
        double u[4]; int aliased = 0;
        if (w->dyn == LSMO_DYN_DI) {
            double a[4] = { g[2], g[3], -g[2], -g[3] };     /* grad @ control_jacobian, safety_filter.py:123-129 */
            if (best_v < eps_hj) {
                for (int k = 0; k < 4; ++k) u[k] = (a[k] < 0.0) ? -0.5 : 0.5;   /* Box.extreme_point (float32) */
            } else {
                double b = g[0] * rel[2] + g[1] * rel[3];
                b = b + p->cbf_rate * best_v;
                const double pinv[4] = { 1.0, 1.0, 1.0, 1.0 };
                if (!qp_project(a, b, uref, pinv, u)) { aliased = 1; for (int k = 0; k < 4; ++k) u[k] = uref[k]; }
            }
            /* clip_ctrl_with_valid_control_bound, safety_filter.py:328-340 (tests RELATIVE velocity, Q3) */
            double dt = p->dt;
            double axmax = (rel[2] < 0.5 - dt * 0.5) ? 0.5 : 0.0;
            double axmin = (rel[2] > -0.5 - dt * (-0.5)) ? -0.5 : 0.0;
            u[0] = pymax(pymin(u[0], axmax), axmin);
            double aymax = (rel[3] < 0.5 - dt * 0.5) ? 0.5 : 0.0;
            double aymin = (rel[3] > -0.5 - dt * (-0.5)) ? -0.5 : 0.0;
            u[1] = pymax(pymin(u[1], aymax), aymin);
        } else {
            const double wmax = 0.1, amin = -0.001, amax = 0.002, vmin = 60 * 0.514444 * 0.001, vmax = 175 * 0.514444 * 0.001;
            /* grad @ control_jacobian (safety_filter.py:53-59), columns [w_a, w_b, a_a, a_b] */
            double a[4];
            a[0] = (g[0] * rel[1] + g[1] * (-rel[0])) + g[2] * (-1.0);
            a[1] = g[2]; a[2] = g[3]; a[3] = g[4];
            int bang = best_v < eps_hj;
            if (bang) {
                /* Air4dCooperativeDynamics.optimal_control_and_disturbance, safety_filter.py:64-78;
                 * boxes are float32 (JAX default dtype) */
                double lo[4] = { f32r(-wmax), f32r(-wmax), f32r(amin), f32r(amin) };
                double hi[4] = { f32r(wmax), f32r(wmax), f32r(amax), f32r(amax) };
                double lo2[4], hi2[4];
                memcpy(lo2, lo, sizeof lo); memcpy(hi2, hi, sizeof hi);
                if (rel[3] <= vmin) { memcpy(lo2, lo, sizeof lo); memcpy(hi2, hi, sizeof hi); lo2[2] = 0.0; }
                if (rel[3] >= vmax) { memcpy(lo2, lo, sizeof lo); memcpy(hi2, hi, sizeof hi); hi2[2] = 0.0; }
                if (rel[4] <= vmin) { memcpy(lo2, lo, sizeof lo); memcpy(hi2, hi, sizeof hi); lo2[3] = 0.0; }
                if (rel[4] >= vmax) { memcpy(lo2, lo, sizeof lo); memcpy(hi2, hi, sizeof hi); hi2[3] = 0.0; }
                for (int k = 0; k < 4; ++k) u[k] = (a[k] < 0.0) ? lo2[k] : hi2[k];
            } else {
                double f0 = -rel[3] + rel[4] * lsm_cos(rel[2]);
                double f1 = rel[4] * lsm_sin(rel[2]);
                double b = g[0] * f0 + g[1] * f1;
                b = b + p->cbf_rate * best_v;
                double pinv[4];
                if (rel[0] < 0.0) { pinv[0] = 1.0 / 100.0; pinv[1] = 1.0 / 10.0; pinv[2] = 1.0 / 10.0; pinv[3] = 1.0 / 1.0; }
                else { pinv[0] = 1.0 / 10.0; pinv[1] = 1.0 / 1.0; pinv[2] = 1.0 / 100.0; pinv[3] = 1.0 / 10.0; }
                if (!qp_project(a, b, uref, pinv, u)) { aliased = 1; for (int k = 0; k < 4; ++k) u[k] = uref[k]; }
                else {
                    u[0] = pymax(pymin(u[0], wmax), -wmax);     /* safety_filter.py:306-307 */
                    u[2] = pymax(pymin(u[2], wmax), -wmax);
                }
            }
            /* clip_ctrl_with_valid_control_bound, safety_filter.py:262-271 */
            double dt = p->dt;
            double cmax = (rel[3] < vmax - dt * amax) ? amax : 0.0;
            double cmin = (rel[3] > vmin - dt * amin) ? amin : 0.0;
            u[1] = pymax(pymin(u[1], cmax), cmin);
            cmax = (rel[4] < vmax - dt * amax) ? amax : 0.0;
            cmin = (rel[4] > vmin - dt * amin) ? amin : 0.0;
            u[3] = pymax(pymin(u[3], cmax), cmin);
            if (bang) { u[1] = f32r(u[1]); u[3] = f32r(u[3]); }   /* stored into a float32 array */
        }
        double nd = 0.0;
        if (!aliased) {
            for (int k = 0; k < 4; ++k) { double d = u[k] - uref[k]; nd = nd + d * d; }
            nd = sqrt(nd);
        }
        filtered[i] = nd > 1e-4;
        safe[i][0] = u[0]; safe[i][1] = u[1];
    }
}

/* core.py:191-210 DoubleIntegratorXYState.update_state (RK45 over one dt == closed form) */
static void integrate_di(env_t *w, int i, const double u[2], double dt) {
    double vx = w->s2[i], vy = w->s3[i];
    w->x[i] = w->x[i] + vx * dt + 0.5 * u[0] * dt * dt;
    w->y[i] = w->y[i] + vy * dt + 0.5 * u[1] * dt * dt;
    vx = vx + u[0] * dt; vy = vy + u[1] * dt;
    double speed = sqrt(vx * vx + vy * vy);
    const double max_speed = 0.5;
    if (speed > max_speed) { vx = max_speed * vx / speed; vy = max_speed * vy / speed; }
    w->s2[i] = vx; w->s3[i] = vy;
    speed = sqrt(vx * vx + vy * vy);
    w->p_dist[i] += speed * dt;
    w->state_time[i] += dt;
}
/* core.py:110-131 KinematicVehicleXYState.update_state; closed form of
 * d[x,y,th,v] = [v cos th, v sin th, w, a] over one dt */
static void integrate_airtaxi(env_t *w, int i, const double u[2], double dt) {
    const double vmin = 60 * 0.514444 * 0.001, vmax = 175 * 0.514444 * 0.001;
    double th0 = w->s2[i], v0 = w->s3[i], om = u[0], ac = u[1];
    double th1 = th0 + om * dt, v1 = v0 + ac * dt;
    double ddx, ddy;
    if (fabs(om * dt) < 1e-3) {
        /* series in (om*dt): integral of (v0 + a t) (cos, sin)(th0 + om t) */
        double T = dt, c0 = lsm_cos(th0), s0 = lsm_sin(th0), o = om;
        double i0 = T, i1 = T * T / 2.0, i2 = T * T * T / 3.0, i3 = T * T * T * T / 4.0, i4 = T * T * T * T * T / 5.0;
        /* cos(th0+ot) ~ c0 - s0 o t - c0 o^2 t^2/2 + s0 o^3 t^3/6 ; sin ~ s0 + c0 o t - s0 o^2 t^2/2 - c0 o^3 t^3/6 */
        double cc0 = c0, cc1 = -s0 * o, cc2 = -c0 * o * o / 2.0, cc3 = s0 * o * o * o / 6.0;
        double sc0 = s0, sc1 = c0 * o, sc2 = -s0 * o * o / 2.0, sc3 = -c0 * o * o * o / 6.0;
        ddx = v0 * (cc0 * i0 + cc1 * i1 + cc2 * i2 + cc3 * i3) + ac * (cc0 * i1 + cc1 * i2 + cc2 * i3 + cc3 * i4);
        ddy = v0 * (sc0 * i0 + sc1 * i1 + sc2 * i2 + sc3 * i3) + ac * (sc0 * i1 + sc1 * i2 + sc2 * i3 + sc3 * i4);
    } else {
        double s1 = lsm_sin(th1), c1 = lsm_cos(th1), s0 = lsm_sin(th0), c0 = lsm_cos(th0);
        ddx = (v1 * s1 - v0 * s0) / om + ac * (c1 - c0) / (om * om);
        ddy = (-(v1 * c1) + v0 * c0) / om + ac * (s1 - s0) / (om * om);
    }
    w->x[i] = w->x[i] + ddx; w->y[i] = w->y[i] + ddy;
    w->s2[i] = th1;
    if (v1 > vmax) v1 = vmax;
    if (v1 < vmin) v1 = vmin;
    w->s3[i] = v1;
    w->p_dist[i] += v1 * dt;
    w->state_time[i] += dt;
}

/* core.py:514-543 calculate_distances: entity order agents, landmarks, obstacles (core.py:489-496) */
static void entity_pos(const env_t *w, int e, double *px, double *py) {
    if (e < w->N) { *px = w->x[e]; *py = w->y[e]; }
    else if (e < w->N + w->M) { *px = w->lx[e - w->N]; *py = w->ly[e - w->N]; }
    else { *px = w->ox[e - w->N - w->M]; *py = w->oy[e - w->N - w->M]; }
}
/* navigation_graph_safe.py:452-465 is_obstacle_collision (no walls): any obstacle closer than 1.05 * (size + size) */
static int obstacle_collision(const env_t *w, double px, double py) {
    for (int k = 0; k < w->O; ++k) {
        double dx = w->ox[k] - px, dy = w->oy[k] - py;
        if (sqrt(dx * dx + dy * dy) < 1.05 * (0.050 + 0.050)) return 1;
    }
    return 0;
}
static void calculate_distances(env_t *w) {
    int E = w->E;
    for (int a = 0; a < E; ++a) {
        w->D[a * E + a] = 0.0;
        double ax, ay; entity_pos(w, a, &ax, &ay);
        for (int b = a + 1; b < E; ++b) {
            double bx, by; entity_pos(w, b, &bx, &by);
            double dx = ax - bx, dy = ay - by;
            double d = sqrt(dx * dx + dy * dy);
            w->D[a * E + b] = d; w->D[b * E + a] = d;
        }
    }
}
/* core.py:696-709 */
static void update_min_relative_distance(env_t *w) {
    for (int i = 0; i < w->N; ++i) {
        double m = INFINITY;
        if (!w->done[i]) for (int j = 0; j < w->N; ++j) {
            if (j == i || w->done[j]) continue;
            double d = norm2(w->x[i] - w->x[j], w->y[i] - w->y[j]);
            if (d < m) m = d;
        }
        w->min_rel[i] = m;
    }
}

/* World.step, core.py:593-631 */
static void world_step(env_t *w, double raw[][2], double safe[][2]) {
    const lsmo_params *p = w->p;
    int world_filter = w->cur[11] != 0.0;
    for (int it = 0; it < p->num_internal_step; ++it) {
        int filtered[MAXN], deconf[MAXN];
        if (world_filter) {
            apply_safety_filter(w, raw, safe, filtered, deconf);
            for (int i = 0; i < w->N; ++i) { w->deconflict[i] = deconf[i]; w->safety_filtered[i] = filtered[i]; }
        } else {
            for (int i = 0; i < w->N; ++i) { safe[i][0] = raw[i][0]; safe[i][1] = raw[i][1]; }
        }
        for (int i = 0; i < w->N; ++i) {
            double d0 = raw[i][0] - safe[i][0], d1 = raw[i][1] - safe[i][1];
            w->action_diff[i] = sqrt(d0 * d0 + d1 * d1);
        }
        for (int i = 0; i < w->N; ++i) {
            if (w->done[i]) continue;
            if (w->dyn == LSMO_DYN_DI) integrate_di(w, i, safe[i], p->dt); else integrate_airtaxi(w, i, safe[i], p->dt);
        }
        calculate_distances(w);
        update_min_relative_distance(w);
    }
}

/* ------------------------------------------------------------------------------------------
 * observation / graph emission
 * ---------------------------------------------------------------------------------------- */
/* navigation_graph_safe.py:855-875 + utils.py:114-137 */
static void emit_obs(const env_t *w, int i, float *out) {
    int g = current_goal(w, i);
    if (w->dyn == LSMO_DYN_DI) {
        out[0] = (float)w->s2[i]; out[1] = (float)w->s3[i];
        out[2] = (float)(w->lx[g] - w->x[i]); out[3] = (float)(w->ly[g] - w->y[i]);
        out[4] = (float)lsm_sin(w->lh[g]); out[5] = (float)lsm_cos(w->lh[g]);
        out[6] = (float)w->ls[g];
    } else {
        double rp[2]; rel_pos_from_reference(w->lx[g], w->ly[g], w->x[i], w->y[i], w->s2[i], rp);
        double rh = w->lh[g] - w->s2[i];
        out[0] = (float)w->s3[i]; out[1] = (float)rp[0]; out[2] = (float)rp[1];
        out[3] = (float)lsm_sin(rh); out[4] = (float)lsm_cos(rh); out[5] = (float)w->ls[g];
    }
}
/* navigation_graph_safe.py:1038-1089 + utils.py:139-255 */
static int node_feat_dim(const env_t *w) {
    if (w->p->flags & LSMO_FLAG_GRAPH_FEAT_GLOBAL) return 7;
    return w->dyn == LSMO_DYN_DI ? 10 : 11;
}

static void emit_node_obs(const env_t *w, int i, float *out) {
    int N = w->N, E = w->E;
    double vi[2]; a_vel(w, i, vi);
    if (w->p->flags & LSMO_FLAG_GRAPH_FEAT_GLOBAL) {
        /* _get_entity_feat_global, navigation_graph_safe.py:1017-1036: [vel, pos, goal_pos, type]; an agent's goal
         * is landmark `optimal_match_index[id]` = its FIRST landmark (:179), a landmark's goal is its own position */
        for (int e = 0; e < E; ++e) {
            float *o = out + (size_t)e * 7;
            if (e < N) {
                double ve[2]; a_vel(w, e, ve);
                o[0] = (float)ve[0]; o[1] = (float)ve[1]; o[2] = (float)w->x[e]; o[3] = (float)w->y[e];
                o[4] = (float)w->lx[e]; o[5] = (float)w->ly[e]; o[6] = 0.0f;
            } else if (e < N + w->M) {
                int m = e - N;
                o[0] = 0.0f; o[1] = 0.0f; o[2] = (float)w->lx[m]; o[3] = (float)w->ly[m];
                o[4] = o[2]; o[5] = o[3]; o[6] = 1.0f;
            } else {    /* obstacle: velocity 0 (state.stop()), goal = own position, type 2 (:1030-1032) */
                int k = e - N - w->M;
                o[0] = 0.0f; o[1] = 0.0f; o[2] = (float)w->ox[k]; o[3] = (float)w->oy[k];
                o[4] = o[2]; o[5] = o[3]; o[6] = 2.0f;
            }
        }
        return;
    }
    for (int e = 0; e < E; ++e) {
        if (w->dyn == LSMO_DYN_DI) {
            float *o = out + (size_t)e * 10;
            if (e < N) {
                int g = current_goal(w, e);
                o[0] = (float)(w->x[e] - w->x[i]); o[1] = (float)(w->y[e] - w->y[i]);
                o[2] = (float)(w->s2[e] - vi[0]); o[3] = (float)(w->s3[e] - vi[1]);
                o[4] = (float)(w->lx[g] - w->x[i]); o[5] = (float)(w->ly[g] - w->y[i]);
                o[6] = (float)lsm_sin(w->lh[g]); o[7] = (float)lsm_cos(w->lh[g]);
                o[8] = (float)w->ls[g]; o[9] = 0.0f;
            } else if (e < N + w->M) {
                int m = e - N;
                o[0] = (float)(w->lx[m] - w->x[i]); o[1] = (float)(w->ly[m] - w->y[i]);
                o[2] = (float)(-vi[0]); o[3] = (float)(-vi[1]);
                o[4] = o[0]; o[5] = o[1];
                o[6] = (float)lsm_sin(w->lh[m]); o[7] = (float)lsm_cos(w->lh[m]);
                o[8] = (float)w->ls[m]; o[9] = 1.0f;
            } else {    /* obstacle (extension): the landmark builder with heading 0, speed 0; type 2 */
                int k = e - N - w->M;
                o[0] = (float)(w->ox[k] - w->x[i]); o[1] = (float)(w->oy[k] - w->y[i]);
                o[2] = (float)(-vi[0]); o[3] = (float)(-vi[1]);
                o[4] = o[0]; o[5] = o[1];
                o[6] = (float)lsm_sin(0.0); o[7] = (float)lsm_cos(0.0);
                o[8] = 0.0f; o[9] = 2.0f;
            }
        } else {
            float *o = out + (size_t)e * 11;
            double thi = w->s2[i];
            if (e < N) {
                int g = current_goal(w, e);
                double rp[2], rg[2], ve[2];
                rel_pos_from_reference(w->x[e], w->y[e], w->x[i], w->y[i], thi, rp);
                double rh = w->s2[e] - thi;
                a_vel(w, e, ve);
                double rs = norm2(ve[0] - vi[0], ve[1] - vi[1]);
                rel_pos_from_reference(w->lx[g], w->ly[g], w->x[i], w->y[i], thi, rg);
                double rgh = w->lh[g] - thi;
                o[0] = (float)rp[0]; o[1] = (float)rp[1]; o[2] = (float)rs;
                o[3] = (float)lsm_sin(rh); o[4] = (float)lsm_cos(rh);
                o[5] = (float)rg[0]; o[6] = (float)rg[1];
                o[7] = (float)lsm_sin(rgh); o[8] = (float)lsm_cos(rgh);
                o[9] = (float)w->ls[g]; o[10] = 0.0f;
            } else {
                /* landmark, or obstacle (extension): the landmark builder with heading 0, speed 0; type 2 */
                int m = e - N, is_obst = e >= N + w->M, k = e - N - w->M;
                double px = is_obst ? w->ox[k] : w->lx[m], py = is_obst ? w->oy[k] : w->ly[m];
                double rp[2];
                rel_pos_from_reference(px, py, w->x[i], w->y[i], thi, rp);
                double rh = (is_obst ? 0.0 : w->lh[m]) - thi;
                o[0] = (float)rp[0]; o[1] = (float)rp[1]; o[2] = (float)w->s3[i];
                o[3] = (float)lsm_sin(rh); o[4] = (float)lsm_cos(rh);
                o[5] = o[0]; o[6] = o[1]; o[7] = o[3]; o[8] = o[4];
                o[9] = is_obst ? 0.0f : (float)w->ls[m]; o[10] = is_obst ? 2.0f : 1.0f;
            }
        }
    }
}
/* navigation_graph_safe.py:974-994: masks the SHARED distance matrix in place (Q1), strict < radius (Q7) */
static void emit_adj(env_t *w, float *out) {
    int N = w->N, E = w->E;
    for (int e = 0; e < E; ++e) {
        int disc;
        if (e < N) disc = w->done[e];
        else if (e < N + w->M) { int m = e - N; disc = w->reached[m % N] > (m / N); }
        else disc = 0;   /* obstacles are never disconnected (extension: the reference's mask has no entry for them) */
        if (disc) for (int k = 0; k < E; ++k) { w->D[e * E + k] = 0.0; w->D[k * E + e] = 0.0; }
    }
    double R = w->p->coordination_range;
    for (int k = 0; k < E * E; ++k) {
        double d = w->D[k];
        out[k] = (d < R && d > 0.0) ? (float)d : 0.0f;
    }
}

/* navigation_graph_safe.py:691-791 */
static double reward_reach_goal(const env_t *w, int i) {
    const lsmo_params *p = w->p;
    double rew = 0.0;
    double sloped = w->cur[0];
    int g = current_goal(w, i);
    double th = a_theta(w, i), sp = a_speed(w, i);
    double he = direction_alignment_error(th, w->lh[g]);
    double hpr = 1.0 - clipd(he / w->cur[2], 0.0, 1.0);
    double se = fabs(sp - w->ls[g]);
    double sen = clipd(se / w->cur[3], 0.0, 1.0);
    int use_filter_arg = (p->flags & LSMO_FLAG_USE_SAFETY_FILTER) != 0;
    double cra = ratio_sloped(w->ratio, 0.25, 0.75);
    if (use_filter_arg) cra = 1.0;
    if (goal_reached(w, i)) {
        double spr = 1.0 - sen;
        /* utils.py:83-89 cross_track_error */
        double pdx = w->lx[g] - w->x[i], pdy = w->ly[g] - w->y[i];
        double cte = pdx * lsm_sin(th) - pdy * lsm_cos(th);
        double nrm = norm2(pdx, pdy);
        cte = fabs(cte) / (nrm > 1e-6 ? nrm : 1e-6);
        cte = clipd(cte, 0.0, 1.0);
        double pr = hpr * spr * (1.0 - cte);
        double goal_rew;
        if (w->dyn == LSMO_DYN_DI) goal_rew = p->goal_rew * pr;
        else goal_rew = p->goal_rew * (pr * cra + (1.0 - cra));
        if (p->flags & LSMO_FLAG_USE_MASKING) { if (!w->done[i]) rew += goal_rew; } else rew += goal_rew;
    }
    if (!w->done[i]) {
        if (w->dyn == LSMO_DYN_DI) {
            if (!use_filter_arg) {
                /* utils.py:323-349 double_integrator_velocity_error_from_magnetic_field_reference */
                double rp[2], rv[2];
                rel_pos_from_reference(w->x[i], w->y[i], w->lx[g], w->ly[g], w->lh[g], rp);
                double dist = norm2(rp[0], rp[1]);
                double ang = lsm_atan2(rp[1], rp[0]);
                const double ang_range = PI / 6;
                rel_pos_from_reference(w->s2[i], w->s3[i], 0.0, 0.0, w->lh[g], rv);
                double rh = lsmo_magnetic_heading(rp[0], rp[1], 2.0 * w->cur[4]);
                double ref_speed = pymax(w->ls[g], 0.1);
                double dr = clipd(dist / 1.5, 0.0, 1.0);
                ref_speed = ref_speed * (1.0 - dr) + 1.0 * dr;
                double ex = rv[0] - ref_speed * lsm_cos(rh), ey = rv[1] - ref_speed * lsm_sin(rh);
                double err = norm2(ex, ey);
                double pen;
                if (lsm_cos(ang) < lsm_cos(ang_range)) pen = err;
                else {
                    double ar = clipd((lsm_cos(ang) - lsm_cos(ang_range)) / (1.0 - lsm_cos(ang_range)), 0.0, 1.0);
                    pen = err * (1.0 - ar) + dist * ar;
                }
                double hap = 3.0 * pen;
                hap = clipd(1.0 - sloped, 0.0, 1.0) * hap;
                rew -= hap;
            }
            if (use_filter_arg) rew -= 1.0; else rew -= 1.0 * sloped;
        } else {
            double rp[2];
            rel_pos_from_reference(w->x[i], w->y[i], w->lx[g], w->ly[g], w->lh[g], rp);
            double rs[4] = { rp[0], rp[1], th - w->lh[g], sp };
            double ttr = lsmo_interpolate(w->tg, rs, -1);
            if (isnan(ttr)) ttr = w->tg->ttr_max;
            rew -= 0.04 * ttr;
            rew -= sen * cra;
        }
    }
    return rew;
}

/* navigation_graph_safe.py:839-853 */
static double reward(env_t *w, int i) {
    const lsmo_params *p = w->p;
    int N = w->N;
    double rew = reward_reach_goal(w, i);
    if (p->flags & LSMO_FLAG_SAFETY_VIOLATION) {              /* :793-798 */
        double r = 0.0;
        for (int a = 0; a < N; ++a) {
            if (a == i) continue;
            double d = norm2(w->x[a] - w->x[i], w->y[a] - w->y[i]);
            if (d < w->cur[9] && !w->done[a]) r += w->cur[6];
        }
        rew += r;
    }
    if (p->flags & LSMO_FLAG_POTENTIAL_CONFLICT) {            /* :800-823 */
        int count = 0; double pen = 0.0;
        double vi[2]; a_vel(w, i, vi);
        for (int a = 0; a < N; ++a) {
            if (a == i) continue;
            double rx = w->x[a] - w->x[i], ry = w->y[a] - w->y[i];
            double rd = norm2(rx, ry);
            if (rd < w->cur[10] && !w->done[a]) {
                double closeness = 1.0 - clipd((rd - w->cur[9]) / (w->cur[10] - w->cur[9]), 0.0, 1.0);
                double dir = lsm_atan2(ry, rx);
                double va[2]; a_vel(w, a, va);
                double change = lsm_cos(dir) * (va[0] - vi[0]) + lsm_sin(dir) * (va[1] - vi[1]);
                change = fabs(pymin(0.0, change));
                pen += change * closeness;
                count += 1;
            }
        }
        if (count > 1) rew += w->cur[5] * pen;
    }
    if ((p->flags & LSMO_FLAG_DIFF_FROM_FILTERED_ACTION) && (p->flags & LSMO_FLAG_USE_SAFETY_FILTER)) {   /* :825-828 */
        if (!w->done[i]) rew += w->cur[7] * w->action_diff[i];
    }
    if (p->flags & LSMO_FLAG_HJ_VALUE) {                      /* :830-837, core.py:459-468 */
        double r = 0.0;
        for (int a = 0; a < N; ++a) {
            if (a == i || w->done[a]) continue;
            double rel[5]; int inr;
            if (w->dyn == LSMO_DYN_DI) di_relative_state(w, i, a, rel); else at_relative_state(w, i, a, rel);
            double v = hj_value(w, rel, &inr);
            double cvp = fabs(pymin(v - 0.4, 0.0));
            r += w->cur[8] * cvp;
        }
        rew += r;
    }
    update_reached_goal_and_done(w, i);
    return clipd(rew, p->min_reward, p->max_reward);
}

/* ------------------------------------------------------------------------------------------
 * load / store between SoA buffers and the working set
 * ---------------------------------------------------------------------------------------- */
#define AF(b, f, e, i, N) ((b)->agent_f64[((size_t)(f) * (size_t)(b)->num_envs + (size_t)(e)) * (size_t)(N) + (size_t)(i)])
#define AI(b, f, e, i, N) ((b)->agent_i32[((size_t)(f) * (size_t)(b)->num_envs + (size_t)(e)) * (size_t)(N) + (size_t)(i)])
#define LF(b, f, e, m, M) ((b)->landmarks[((size_t)(f) * (size_t)(b)->num_envs + (size_t)(e)) * (size_t)(M) + (size_t)(m)])
#define EF(b, f, e) ((b)->env_f64[(size_t)(f) * (size_t)(b)->num_envs + (size_t)(e)])
#define EI(b, f, e) ((b)->env_i32[(size_t)(f) * (size_t)(b)->num_envs + (size_t)(e)])

static void env_load(env_t *w, const lsmo_params *p, const lsmo_grid *vg, const lsmo_grid *tg,
                     const lsmo_buffers *b, int64_t e) {
    w->p = p; w->vg = vg; w->tg = tg;
    int N = w->N = p->num_agents; w->L = p->num_landmarks; int M = w->M = N * p->num_landmarks;
    int O = w->O = p->num_obstacles;
    w->E = N + M + O; w->dyn = p->dynamics;
    w->current_step = EI(b, LSMO_EI_CURRENT_STEP, e);
    w->reset_count = EI(b, LSMO_EI_RESET_COUNT, e);
    w->parity = EI(b, LSMO_EI_PARITY, e);
    w->ratio = EF(b, LSMO_EF_CURRICULUM_RATIO, e);
    lsmo_curriculum(p, w->ratio, w->cur);
    int told = w->parity ? LSMO_AF_TIMES_REQ_B : LSMO_AF_TIMES_REQ_A;
    int dold = w->parity ? LSMO_AF_DISTS_GOAL_B : LSMO_AF_DISTS_GOAL_A;
    for (int i = 0; i < N; ++i) {
        w->x[i] = AF(b, LSMO_AF_X, e, i, N); w->y[i] = AF(b, LSMO_AF_Y, e, i, N);
        w->s2[i] = AF(b, LSMO_AF_S2, e, i, N); w->s3[i] = AF(b, LSMO_AF_S3, e, i, N);
        w->p_dist[i] = AF(b, LSMO_AF_P_DIST, e, i, N); w->state_time[i] = AF(b, LSMO_AF_STATE_TIME, e, i, N);
        w->min_rel[i] = AF(b, LSMO_AF_MIN_REL_DIST, e, i, N); w->goal_min_time[i] = AF(b, LSMO_AF_GOAL_MIN_TIME, e, i, N);
        w->times_req_old[i] = AF(b, told, e, i, N); w->dists_goal_old[i] = AF(b, dold, e, i, N);
        w->times_req[i] = w->times_req_old[i]; w->dists_goal[i] = w->dists_goal_old[i];
        w->dist_left[i] = AF(b, LSMO_AF_DIST_LEFT, e, i, N);
        w->ep_travel_dist[i] = AF(b, LSMO_AF_EP_TRAVEL_DIST, e, i, N); w->ep_min_dist[i] = AF(b, LSMO_AF_EP_MIN_DIST, e, i, N);
        w->action_diff[i] = AF(b, LSMO_AF_ACTION_DIFF, e, i, N);
        w->reached[i] = AI(b, LSMO_AI_REACHED, e, i, N); w->done[i] = AI(b, LSMO_AI_DONE, e, i, N);
        w->safety_filtered[i] = AI(b, LSMO_AI_SAFETY_FILTERED, e, i, N); w->deconflict[i] = AI(b, LSMO_AI_DECONFLICT_IDX, e, i, N);
        w->ncoll[i] = AI(b, LSMO_AI_NUM_COLLISIONS, e, i, N); w->ep_travel_len[i] = AI(b, LSMO_AI_EP_TRAVEL_LEN, e, i, N);
        w->ep_conflict[i] = AI(b, LSMO_AI_EP_CONFLICT, e, i, N); w->ep_multi[i] = AI(b, LSMO_AI_EP_MULTI, e, i, N);
        w->ep_done[i] = AI(b, LSMO_AI_EP_DONE, e, i, N);
        w->nobst[i] = AI(b, LSMO_AI_NUM_OBST_COLLISIONS, e, i, N);
    }
    for (int k = 0; k < O; ++k) {
        w->ox[k] = b->obstacles[((size_t)0 * (size_t)b->num_envs + (size_t)e) * (size_t)O + k];
        w->oy[k] = b->obstacles[((size_t)1 * (size_t)b->num_envs + (size_t)e) * (size_t)O + k];
    }
    for (int m = 0; m < M; ++m) {
        w->lx[m] = LF(b, LSMO_LF_X, e, m, M); w->ly[m] = LF(b, LSMO_LF_Y, e, m, M);
        w->lh[m] = LF(b, LSMO_LF_HEADING, e, m, M); w->ls[m] = LF(b, LSMO_LF_SPEED, e, m, M);
        w->lsin[m] = LF(b, LSMO_LF_SIN, e, m, M); w->lcos[m] = LF(b, LSMO_LF_COS, e, m, M);
    }
}

static void env_store(const env_t *w, const lsmo_buffers *b, int64_t e, int store_landmarks) {
    int N = w->N, M = w->M;
    EI(b, LSMO_EI_CURRENT_STEP, e) = w->current_step;
    EI(b, LSMO_EI_RESET_COUNT, e) = w->reset_count;
    EI(b, LSMO_EI_PARITY, e) = w->parity;
    EF(b, LSMO_EF_CURRICULUM_RATIO, e) = w->ratio;
    /* `parity` already names the slot holding the NEWEST values */
    int tnew = w->parity ? LSMO_AF_TIMES_REQ_B : LSMO_AF_TIMES_REQ_A;
    int dnew = w->parity ? LSMO_AF_DISTS_GOAL_B : LSMO_AF_DISTS_GOAL_A;
    for (int i = 0; i < N; ++i) {
        AF(b, LSMO_AF_X, e, i, N) = w->x[i]; AF(b, LSMO_AF_Y, e, i, N) = w->y[i];
        AF(b, LSMO_AF_S2, e, i, N) = w->s2[i]; AF(b, LSMO_AF_S3, e, i, N) = w->s3[i];
        AF(b, LSMO_AF_P_DIST, e, i, N) = w->p_dist[i]; AF(b, LSMO_AF_STATE_TIME, e, i, N) = w->state_time[i];
        AF(b, LSMO_AF_MIN_REL_DIST, e, i, N) = w->min_rel[i]; AF(b, LSMO_AF_GOAL_MIN_TIME, e, i, N) = w->goal_min_time[i];
        AF(b, tnew, e, i, N) = w->times_req[i]; AF(b, dnew, e, i, N) = w->dists_goal[i];
        AF(b, LSMO_AF_DIST_LEFT, e, i, N) = w->dist_left[i];
        AF(b, LSMO_AF_EP_TRAVEL_DIST, e, i, N) = w->ep_travel_dist[i]; AF(b, LSMO_AF_EP_MIN_DIST, e, i, N) = w->ep_min_dist[i];
        AF(b, LSMO_AF_ACTION_DIFF, e, i, N) = w->action_diff[i];
        AI(b, LSMO_AI_REACHED, e, i, N) = w->reached[i]; AI(b, LSMO_AI_DONE, e, i, N) = w->done[i];
        AI(b, LSMO_AI_SAFETY_FILTERED, e, i, N) = w->safety_filtered[i]; AI(b, LSMO_AI_DECONFLICT_IDX, e, i, N) = w->deconflict[i];
        AI(b, LSMO_AI_NUM_COLLISIONS, e, i, N) = w->ncoll[i]; AI(b, LSMO_AI_EP_TRAVEL_LEN, e, i, N) = w->ep_travel_len[i];
        AI(b, LSMO_AI_EP_CONFLICT, e, i, N) = w->ep_conflict[i]; AI(b, LSMO_AI_EP_MULTI, e, i, N) = w->ep_multi[i];
        AI(b, LSMO_AI_EP_DONE, e, i, N) = w->ep_done[i];
        AI(b, LSMO_AI_NUM_OBST_COLLISIONS, e, i, N) = w->nobst[i];
    }
    if (store_landmarks) for (int k = 0; k < w->O; ++k) {
        b->obstacles[((size_t)0 * (size_t)b->num_envs + (size_t)e) * (size_t)w->O + k] = w->ox[k];
        b->obstacles[((size_t)1 * (size_t)b->num_envs + (size_t)e) * (size_t)w->O + k] = w->oy[k];
    }
    if (store_landmarks) for (int m = 0; m < M; ++m) {
        LF(b, LSMO_LF_X, e, m, M) = w->lx[m]; LF(b, LSMO_LF_Y, e, m, M) = w->ly[m];
        LF(b, LSMO_LF_HEADING, e, m, M) = w->lh[m]; LF(b, LSMO_LF_SPEED, e, m, M) = w->ls[m];
        LF(b, LSMO_LF_SIN, e, m, M) = w->lsin[m]; LF(b, LSMO_LF_COS, e, m, M) = w->lcos[m];
    }
}

/* ------------------------------------------------------------------------------------------
 * reset: environment.py:1046-1074, navigation_graph_safe.py:264-317, :1199-1367, utils.py:39-68
 * ---------------------------------------------------------------------------------------- */
static void sample_separated_positions(rng_t *r, int num, double xlo, double xhi, double ylo, double yhi,
                                       double min_d, double max_d, double px[], double py[]) {
    for (int i = 0; i < num; ++i) {
        double x = 0.0, y = 0.0;
        if (i > 0) {
            for (int j = 0; j < 1000; ++j) {
                x = rng_uniform(r, xlo, xhi); y = rng_uniform(r, ylo, yhi);
                double dm = INFINITY;
                for (int k = 0; k < i; ++k) { double d = norm2(px[k] - x, py[k] - y); if (d < dm) dm = d; }
                if (dm > min_d && dm < max_d) break;
            }
        } else { x = rng_uniform(r, xlo, xhi); y = rng_uniform(r, ylo, yhi); }
        px[i] = x; py[i] = y;
    }
}

static void random_scenario(env_t *w, rng_t *r) {
    const lsmo_params *p = w->p;
    int N = w->N, L = w->L;
    double ws = p->world_size;
    int use_filter_arg = (p->flags & LSMO_FLAG_USE_SAFETY_FILTER) != 0;
    double cra = ratio_sloped(w->ratio, 0.25, 0.75);
    if (use_filter_arg) cra = 1.0;
    /* static obstacles first (:1204-1209): 0.8 * uniform(-ws/2, ws/2, 2) */
    for (int k = 0; k < w->O; ++k) {
        w->ox[k] = 0.8 * rng_uniform(r, -ws / 2.0, ws / 2.0);
        w->oy[k] = 0.8 * rng_uniform(r, -ws / 2.0, ws / 2.0);
    }
    for (int i = 0; i < N; ++i) {
        /* :1218-1249: redraw the position while it collides with an obstacle (the reference loops without a bound; 1000
         * tries here); the airtaxi speed / heading are drawn once the position is accepted */
        for (int tries = 0; tries < 1000; ++tries) {
            if (w->dyn == LSMO_DYN_DI) {
                w->x[i] = rng_uniform(r, -0.8 * ws, 0.8 * ws);
                w->y[i] = rng_uniform(r, -0.8 * ws, 0.8 * ws);
            } else {
                double xmin = -0.5 * ws;
                double xmax = 0.25 * ws * cra + 0.0 * (1.0 - cra) * ws;
                double ry = rng_uniform(r, -0.5 * ws, 0.5 * ws);
                w->x[i] = rng_uniform(r, xmin, xmax); w->y[i] = ry;
            }
            if (!obstacle_collision(w, w->x[i], w->y[i])) break;
        }
        if (w->dyn == LSMO_DYN_DI) { w->s2[i] = 0.0; w->s3[i] = 0.0; }
        else {
            double speed = rng_uniform(r, p->goal_speed_min, p->goal_speed_max);
            w->s2[i] = rng_uniform(r, 0.0, 2.0 * PI);
            w->s3[i] = speed;
        }
        w->done[i] = 0;
    }
    double prevx[MAXM], prevy[MAXM]; int have_prev = 0;
    for (int i = 0; i < N; ++i) {
        double gx[MAXM], gy[MAXM], gh[MAXM] = { 0.0 }, gs[MAXM];
        if (w->dyn == LSMO_DYN_DI) {
            sample_separated_positions(r, L, -0.5 * ws, 0.5 * ws, -0.5 * ws, 0.5 * ws,
                                       0.25 * p->coordination_range, 0.75 * p->coordination_range, gx, gy);
            if (have_prev) for (int l = 0; l < L; ++l) if (rng_uniform(r, 0.0, 1.0) < 0.5) { gx[l] = prevx[l]; gy[l] = prevy[l]; }
        } else {
            double yw = 0.1 * (1.0 - cra) + 0.5 * cra;
            sample_separated_positions(r, L, 0.0, 0.75 * ws, -yw * ws, yw * ws,
                                       0.5 * p->coordination_range, p->coordination_range, gx, gy);
            if (have_prev) for (int l = 0; l < L; ++l) if (rng_uniform(r, 0.0, 1.0) < 0.5) { gx[l] = prevx[l]; gy[l] = prevy[l]; }
            if (gx[0] > gx[1]) { double tx = gx[0], ty = gy[0]; gx[0] = gx[1]; gy[0] = gy[1]; gx[1] = tx; gy[1] = ty; }
        }
        for (int l = 0; l < L - 1; ++l) gh[l] = lsm_atan2(gy[l + 1] - gy[l], gx[l + 1] - gx[l]);
        double last_heading = gh[L - 2];
        double cr = use_filter_arg ? 1.0 : ratio_sloped(w->ratio, 0.25, 0.75);
        if (w->dyn == LSMO_DYN_AIRTAXI) {
            for (int l = 0; l < L; ++l) gs[l] = p->goal_speed_max;
        } else {
            double rnd[MAXM];
            for (int l = 0; l < L; ++l) rnd[l] = rng_uniform(r, p->goal_speed_min, p->goal_speed_max);
            double var = rng_uniform(r, 0.0, 1.0);
            if (var < pymin(cr, 1.0 - 0.2)) for (int l = 0; l < L; ++l) gs[l] = rnd[l];
            else { for (int l = 0; l < L; ++l) gs[l] = p->goal_speed_max; gs[L - 1] = p->goal_speed_min; }
        }
        for (int l = 0; l < L - 1; ++l) {
            double pr = (w->dyn == LSMO_DYN_DI) ? cr * 0.25 * PI : cra * 0.1 * PI;
            gh[l] += rng_uniform(r, -pr, pr);
        }
        gh[L - 1] = last_heading;
        for (int l = 0; l < L; ++l) {
            int m = l * N + i;
            w->lx[m] = gx[l]; w->ly[m] = gy[l]; w->lh[m] = gh[l]; w->ls[m] = gs[l];
            w->lsin[m] = lsm_sin(gh[l]); w->lcos[m] = lsm_cos(gh[l]);
            prevx[l] = gx[l]; prevy[l] = gy[l];
        }
        have_prev = 1;
    }
}

/* environment.py:895-926 save_summary_of_episode, with episode_agent_reached_goals_list taken
 * from the scenario BEFORE the world reset (environment.py:1047-1048) */
static void episode_summary(const env_t *w, double out[LSMO_EP_COUNT]) {
    int N = w->N;
    double dt = w->p->dt;
    double s_len = 0, s_dist = 0, s_done = 0, s_reached = 0, s_conf = 0, s_min = 0, s_multi = 0, mn = INFINITY;
    for (int i = 0; i < N; ++i) {
        s_len += (double)w->ep_travel_len[i]; s_dist += w->ep_travel_dist[i];
        s_done += (double)w->ep_done[i]; s_reached += (double)w->reached[i];
    }
    for (int i = 0; i < N; ++i) {
        double len = w->ep_travel_len[i] == 0 ? 1.0 : (double)w->ep_travel_len[i];
        s_conf += (double)w->ep_conflict[i] / len;
        s_multi += (double)w->ep_multi[i] / len;
        s_min += w->ep_min_dist[i];
        if (w->ep_min_dist[i] < mn) mn = w->ep_min_dist[i];
    }
    out[LSMO_EP_TRAVEL_TIME_MEAN] = dt * (s_len / N);
    out[LSMO_EP_TRAVEL_DISTANCE_MEAN] = s_dist / N;
    out[LSMO_EP_DONE_PERCENTAGE] = s_done / N;
    out[LSMO_EP_NUM_REACHED_GOAL_MEAN] = s_reached / N;
    out[LSMO_EP_CONFLICT_PERCENTAGE] = s_conf / N;
    double mm = s_min / N;
    out[LSMO_EP_MIN_DISTANCE_MEAN] = isinf(mm) ? w->p->coordination_range : mm;
    out[LSMO_EP_MIN_DISTANCE_MIN] = isinf(mn) ? w->p->coordination_range : mn;
    out[LSMO_EP_MULTIPLE_ENGAGEMENT_PERCENTAGE] = s_multi / N;
}

static void emit_all(env_t *w, const lsmo_buffers *b, int64_t e) {
    int N = w->N, E = w->E;
    int D = w->dyn == LSMO_DYN_DI ? 7 : 6, F = node_feat_dim(w);
    calculate_distances(w);
    for (int i = 0; i < N; ++i) {
        emit_obs(w, i, b->obs + ((size_t)e * N + i) * D);
        emit_node_obs(w, i, b->node_obs + ((size_t)e * N + i) * (size_t)E * F);
        emit_adj(w, b->adj + ((size_t)e * N + i) * (size_t)E * E);
    }
}

static void env_reset(env_t *w, const lsmo_buffers *b, int64_t e, int64_t episode, uint64_t seed, int sample) {
    const lsmo_params *p = w->p;
    int N = w->N;
    double summary[LSMO_EP_COUNT];
    episode_summary(w, summary);
    for (int k = 0; k < LSMO_EP_COUNT; ++k) b->ep_info[(size_t)e * LSMO_EP_COUNT + k] = summary[k];
    w->current_step = 0;
    /* update_curriculum, navigation_graph_safe.py:326 */
    w->ratio = clipd((double)episode / (double)p->num_total_episode, 0.0, 1.0);
    lsmo_curriculum(p, w->ratio, w->cur);
    if (sample) {
        rng_t r; rng_init(&r, seed, (uint32_t)(b->env_id_base + e), (uint32_t)w->reset_count);
        random_scenario(w, &r);
        w->reset_count += 1;
    } else {
        for (int i = 0; i < N; ++i) w->done[i] = 0;
    }
    for (int i = 0; i < N; ++i) {
        w->p_dist[i] = 0.0; w->state_time[i] = 0.0;
        /* min_time, navigation_graph_safe.py:525-535 (landmark id == agent id) */
        w->goal_min_time[i] = norm2(w->x[i] - w->lx[i], w->y[i] - w->ly[i]) / p->agent_max_speed;
        w->reached[i] = 0;
        w->times_req[i] = -1.0; w->times_req_old[i] = -1.0;
        w->dists_goal[i] = -1.0; w->dists_goal_old[i] = -1.0; w->dist_left[i] = -1.0;
        w->ncoll[i] = 0; w->nobst[i] = 0;
        w->ep_travel_len[i] = 0; w->ep_travel_dist[i] = 0.0; w->ep_done[i] = 0;
        w->ep_conflict[i] = 0; w->ep_multi[i] = 0; w->ep_min_dist[i] = INFINITY;
    }
    emit_all(w, b, e);
}

/* ------------------------------------------------------------------------------------------
 * one env.step: multiagent/environment.py:963-1042 (+ graphworker auto reset, env_wrappers.py:851-875)
 * ---------------------------------------------------------------------------------------- */
static void env_step(env_t *w, const lsmo_buffers *b, int64_t e, const int32_t *act, int64_t episode,
                     uint64_t seed, int auto_reset) {
    const lsmo_params *p = w->p;
    int N = w->N, E = w->E;
    int D = w->dyn == LSMO_DYN_DI ? 7 : 6, F = node_feat_dim(w);
    w->current_step += 1;
    double raw[MAXN][2], safe[MAXN][2];
    for (int i = 0; i < N; ++i) {          /* _set_action, environment.py:387-410 */
        int idx = act[i];
        int i0 = idx / 5, i1 = idx - i0 * 5;
        raw[i][0] = p->act_tab0[i0]; raw[i][1] = p->act_tab1[i1];
    }
    world_step(w, raw, safe);
    for (int i = 0; i < N; ++i) { b->safe_action[((size_t)e * N + i) * 2] = safe[i][0]; b->safe_action[((size_t)e * N + i) * 2 + 1] = safe[i][1]; }
    double rew[MAXN]; int all_done = 1;
    for (int i = 0; i < N; ++i) {
        emit_obs(w, i, b->obs + ((size_t)e * N + i) * D);
        rew[i] = reward(w, i);
        float *adj_i = b->adj + ((size_t)e * N + i) * (size_t)E * E;
        emit_node_obs(w, i, b->node_obs + ((size_t)e * N + i) * (size_t)E * F);
        emit_adj(w, adj_i);
        /* episode statistics, environment.py:1004-1022 (reads the float64 masked matrix) */
        if (!w->done[i]) {
            w->ep_travel_len[i] += 1;
            double v[2]; a_vel(w, i, v);
            w->ep_travel_dist[i] += norm2(v[0], v[1]) * p->dt;
            int cnt = 0, have = 0; double mn = INFINITY;
            double R = p->coordination_range;
            for (int j = 0; j < N; ++j) {
                double d = w->D[i * E + j];
                d = (d < R && d > 0.0) ? d : 0.0;
                if (d != 0.0) { have = 1; if (d < p->engagement_distance_ref) cnt++; if (d < mn) mn = d; }
            }
            if (have) {
                if (cnt > 1) w->ep_multi[i] += 1;
                if (mn < p->separation_distance_target) w->ep_conflict[i] += 1;
                if (mn < w->ep_min_dist[i]) w->ep_min_dist[i] = mn;
            }
        }
        if (w->done[i]) w->ep_done[i] = 1;
        int done_out = w->done[i] || (w->current_step >= p->episode_length);   /* environment.py:260-268 */
        b->done[(size_t)e * N + i] = (uint8_t)done_out;
        if (!done_out) all_done = 0;
        /* info_callback state, navigation_graph_safe.py:386-413 */
        {
            int g = current_goal(w, i);
            double dx = w->x[i] - w->lx[g], dy = w->y[i] - w->ly[g];
            double dist = sqrt(dx * dx + dy * dy);
            if (goal_reached(w, i) && w->times_req[i] == -1.0) {
                w->times_req[i] = (double)w->current_step * p->dt;
                w->dists_goal[i] = w->p_dist[i];
                w->dist_left[i] = dist;
            }
            if (w->times_req[i] == -1.0) { w->dists_goal[i] = w->p_dist[i]; w->dist_left[i] = dist; }
            if (obstacle_collision(w, w->x[i], w->y[i])) w->nobst[i] += 1;     /* :402-404 */
            for (int a = 0; a < N; ++a) {
                if (a == i) continue;
                double d = norm2(w->x[i] - w->x[a], w->y[i] - w->y[a]);
                if (d < 1.05 * (0.050 + 0.050)) w->ncoll[i] += 1;
            }
        }
    }
    if (p->flags & LSMO_FLAG_SHARED_REWARD) {       /* environment.py:1032-1037 */
        double s = 0.0; for (int i = 0; i < N; ++i) s += rew[i];
        for (int i = 0; i < N; ++i) rew[i] = s;
    }
    for (int i = 0; i < N; ++i) b->reward[(size_t)e * N + i] = (float)rew[i];
    w->parity ^= 1;     /* the slot written by env_store now holds this step's times_required / dists_to_goal */
    int just_reset = 0;
    if (auto_reset && all_done) { env_reset(w, b, e, episode, seed, 1); just_reset = 1; }
    EI(b, LSMO_EI_JUST_RESET, e) = just_reset;
}

/* ------------------------------------------------------------------------------------------
 * exported entry points
 * ---------------------------------------------------------------------------------------- */
static int check(const lsmo_params *p) {
    if (p->num_agents < 1 || p->num_agents > MAXN) return 1;
    if (p->num_landmarks < 2 || p->num_agents * p->num_landmarks > MAXM) return 1;
    if (p->dynamics != LSMO_DYN_DI && p->dynamics != LSMO_DYN_AIRTAXI) return 1;
    if (p->num_obstacles < 0 || p->num_obstacles > MAXO) return 1;
    return 0;
}

typedef struct {
    int mode;   /* 0 step, 1 reset, 2 observe */
    const lsmo_params *p; const lsmo_grid *vg, *tg; const lsmo_buffers *b;
    const int32_t *action_idx; const uint8_t *env_mask;
    int64_t episode; uint64_t seed; int flag;   /* flag: auto_reset (step) | sample (reset) */
    int64_t e0, e1;
} job_t;

static void *job_run(void *arg) {
    job_t *j = (job_t *)arg;
    const lsmo_params *p = j->p; const lsmo_buffers *b = j->b;
    int N = p->num_agents, E = N * (1 + p->num_landmarks) + p->num_obstacles;
    env_t *w = (env_t *)malloc(sizeof(env_t));
    w->D = (double *)malloc(sizeof(double) * (size_t)E * E);
    for (int64_t e = j->e0; e < j->e1; ++e) {
        if (j->mode == 0) {
            env_load(w, p, j->vg, j->tg, b, e);
            env_step(w, b, e, j->action_idx + (size_t)e * N, j->episode, j->seed, j->flag);
            int just_reset = EI(b, LSMO_EI_JUST_RESET, e);
            if (just_reset) {   /* after a reset both ping-pong slots hold the reset value */
                int saved = w->parity;
                w->parity ^= 1; env_store(w, b, e, 1);
                w->parity = saved;
            }
            env_store(w, b, e, just_reset);
        } else if (j->mode == 1) {
            if (j->env_mask && !j->env_mask[e]) continue;
            env_load(w, p, j->vg, j->tg, b, e);
            env_reset(w, b, e, j->episode, j->seed, j->flag);
            int saved = w->parity;
            w->parity ^= 1; env_store(w, b, e, 1);
            w->parity = saved; env_store(w, b, e, 1);
            EI(b, LSMO_EI_JUST_RESET, e) = 1;
        } else {
            env_load(w, p, NULL, NULL, b, e);
            emit_all(w, b, e);
        }
    }
    free(w->D); free(w);
    return NULL;
}

static int run_jobs(job_t *proto, int nthreads) {
    int64_t n = proto->b->num_envs;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if ((int64_t)nthreads > n) nthreads = (int)(n > 0 ? n : 1);
    if (nthreads == 1) { proto->e0 = 0; proto->e1 = n; job_run(proto); return 0; }
    pthread_t th[256]; job_t jobs[256];
    for (int t = 0; t < nthreads; ++t) {
        jobs[t] = *proto;
        jobs[t].e0 = n * t / nthreads; jobs[t].e1 = n * (t + 1) / nthreads;
        if (pthread_create(&th[t], NULL, job_run, &jobs[t]) != 0) return 2;
    }
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    return 0;
}

int lsmo_step(const lsmo_params *p, const lsmo_grid *vg, const lsmo_grid *tg, const lsmo_buffers *b,
              const int32_t *action_idx, int64_t episode, uint64_t seed, int auto_reset, int nthreads) {
    g_interp_f32 = (p->flags & LSMO_FLAG_INTERP_FLOAT32) != 0;
    if (check(p)) return 1;
    job_t j; memset(&j, 0, sizeof j);
    j.mode = 0; j.p = p; j.vg = vg; j.tg = tg; j.b = b; j.action_idx = action_idx;
    j.episode = episode; j.seed = seed; j.flag = auto_reset;
    return run_jobs(&j, nthreads);
}

int lsmo_reset(const lsmo_params *p, const lsmo_grid *vg, const lsmo_grid *tg, const lsmo_buffers *b,
               const uint8_t *env_mask, int64_t episode, uint64_t seed, int sample, int nthreads) {
    g_interp_f32 = (p->flags & LSMO_FLAG_INTERP_FLOAT32) != 0;
    if (check(p)) return 1;
    job_t j; memset(&j, 0, sizeof j);
    j.mode = 1; j.p = p; j.vg = vg; j.tg = tg; j.b = b; j.env_mask = env_mask;
    j.episode = episode; j.seed = seed; j.flag = sample;
    return run_jobs(&j, nthreads);
}

int lsmo_observe(const lsmo_params *p, const lsmo_buffers *b, int nthreads) {
    g_interp_f32 = (p->flags & LSMO_FLAG_INTERP_FLOAT32) != 0;
    if (check(p)) return 1;
    job_t j; memset(&j, 0, sizeof j);
    j.mode = 2; j.p = p; j.b = b;
    return run_jobs(&j, nthreads);
}

/* the shared math of include/lsm_math.h, exported for tests (op: 0 sin, 1 cos, 2 atan2(a, b)) */
void lsmo_math_eval(int op, const double *a, const double *b, double *out, int64_t n) {
    for (int64_t k = 0; k < n; ++k)
        out[k] = op == 0 ? lsm_sin(a[k]) : (op == 1 ? lsm_cos(a[k]) : lsm_atan2(a[k], b[k]));
}
