// lsm_kernel_spec.cuh - step kernel specialised at compile time on (dynamics, N agents, L landmarks
// per agent): constant-size shared-memory records, unrolled loops, immediate-offset stores.
//
// Same decisions as the generic kernel for every thresholded quantity, organised for far fewer instructions:
//   * envs per warp (EPW) is a LAUNCH choice (1 .. 32/G): small batches spread over more warps so the
//     whole batch is resident in one wave; large batches pack 32/G environments per warp.
//   * the HJ value lookups run pair-parallel: every (ego, other) pair of the warp's environments is one
//     lane-task, results go to shared memory, the agent lane then takes the first minimum in `other`
//     order exactly like np.argmin. The stencil uses 32-bit indices and an exactly rounded division by
//     the grid spacing through its precomputed reciprocal (Markstein: q0 = a*y, r = fma(-b,q0,a),
//     q = fma(r,y,q0) equals RN(a/b); checked on the CPU in tests/test_host_logic.py).
//   * distances are kept SQUARED in float64; every `d < T` / `d > T` test of the reference becomes
//     `d2 < T2` against a host-computed exact squared threshold (sqrt_rn is monotone, so
//     {t : sqrt_rn(t) >= T} is an interval whose lower end the host finds with nextafter). Square roots
//     are taken only where a distance VALUE is stored (min distance, float32 adjacency).
//   * the float32 adjacency value is d2f * rsqrt(d2f) (<= 4e-7 relative from the float64 reference, bar
//     1e-5); which entries are non-zero is decided exactly in float64.
//   * node features are branch-free rows over unified per-entity tables (landmarks carry zero velocity and
//     their own position as "goal"), read with 16-byte shared-memory loads.
#pragma once
#include "lsm_step_common.cuh"

namespace lsm {

template <int DYN, int N, int L>
struct __align__(16) EnvShared {
    static constexpr int M = N * L;
    static constexpr int E = N + M;
    static constexpr int W = (E + 31) / 32;
    // unified per-entity tables for the branch-free node-feature rows:
    //   pos[e]                     position of entity e (agents after the dynamics, then landmarks)
    //   pos[E + s*N + a]           goal position of agent a before (s=0) / after (s=1) its own goal update
    //   vel[s*N + a], vel[2N] = 0  world-frame velocity of agent a before / after its update; landmarks use slot 2N
    //   cst[m], cst[M + s*N + a]   (sin heading, cos heading, speed, type) of landmark m / of agent a's goal
    double2 pos[E + 2 * N];
    double2 vel[2 * N + 1];
    float4 cst[M + 2 * N];
    double as2[N], as3[N];     // state components 2,3 BEFORE the own goal update (vx,vy | theta,speed)
    double rawx[N], rawy[N];   // decoded raw controls
    double sth[N], cth[N], spd_post[N];       // airtaxi: sin/cos(theta), speed after the own update
    double lh[M], lsp[M], lsin[M], lcos[M];   // landmark heading, speed, sin/cos(heading)
    union alignas(16) {
        float dthr[E * E];     // P4: radius-thresholded distance matrix (float32, what adj stores)
        struct {
            double d2aa[N * N];    // P1/P2: squared agent-agent distances (before / after the dynamics)
            double fval[N * N];    // P1: HJ value of (ego i, other j); +inf = out of range
        };
    };
    int goal[2][N], reached[2][N], done[2][N];
    unsigned disc[2][W], keepm[N * W];
    double cur_sep;            // scenario.separation_distance of this env (curriculum)
    int cur_filter;            // world.use_safety_filter of this env (curriculum, Q5)
};

template <int N> struct Pow2 { static constexpr int value = N <= 1 ? 1 : N <= 2 ? 2 : N <= 4 ? 4 : N <= 8 ? 8 : N <= 16 ? 16 : 32; };

// ---------------------------------------------------------------------------------------------
// lean stencil: 32-bit indices, exact division through the reciprocal, weight prefixes shared
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double div_exact(double a, double b, double y /* RN(1/b) */) {
    const double q0 = a * y;
    const double r = fma(-b, q0, a);
    return fma(r, y, q0);
}

template <int ND>
struct Stencil32 {
    int lo[ND], hi[ND];        // linear offset contributions
    double wlo[ND], whi[ND];
    bool valid;
};

template <int ND>
__device__ __forceinline__ void stencil32_setup(const GridDev& g, const double (&x)[ND], Stencil32<ND>& s) {
    s.valid = true;
    int mul = 1;
#pragma unroll
    for (int d = ND - 1; d >= 0; --d) {
        double pos = div_exact(x[d] - g.lo[d], g.spacing[d], g.inv_spacing[d]);
        if (isnan(pos)) s.valid = false;
        if (!(fabs(pos) <= 1.0e9)) pos = pos < 0.0 ? -1.0e9 : 1.0e9;     // declared clamp; never taken in-grid
        const double fl = floor(pos);
        const double whi = pos - fl;
        s.wlo[d] = 1.0 - whi; s.whi[d] = whi;
        const int n = g.shape[d];
        int il = (int)fl, ih = il + 1;          // |fl| <= 1e9 fits in int32
        if (g.periodic[d]) {
            il %= n; if (il < 0) il += n;
            ih %= n; if (ih < 0) ih += n;
        } else {
            il = min(max(il, 0), n - 1);
            ih = min(max(ih, 0), n - 1);
        }
        s.lo[d] = il * mul; s.hi[d] = ih * mul;
        mul *= n;
    }
}

// value: sum over corners (binary counting, dim 0 slowest) of ((w0*w1)*w2..)*v, sequential adds.
// All 2^ND loads are issued first (one memory round trip per lookup instead of 2^ND dependent ones).
template <int ND>
__device__ __forceinline__ double stencil32_value(const GridDev& g, const Stencil32<ND>& s) {
    constexpr int NC = 1 << ND;
    float v[NC];
#pragma unroll
    for (int corner = 0; corner < NC; ++corner) {
        int lin = 0;
#pragma unroll
        for (int d = 0; d < ND; ++d) lin += ((corner >> (ND - 1 - d)) & 1) ? s.hi[d] : s.lo[d];
        v[corner] = __ldg(g.values + lin);
    }
    double acc = 0.0;
#pragma unroll
    for (int corner = 0; corner < NC; ++corner) {
        double weight = 0.0;
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            const double wd = ((corner >> (ND - 1 - d)) & 1) ? s.whi[d] : s.wlo[d];
            weight = (d == 0) ? wd : weight * wd;
        }
        acc = acc + weight * (double)v[corner];
    }
    return acc;
}

// all gradient components in one pass over the corners (same corner / weight order as the value)
template <int ND>
__device__ __forceinline__ void stencil32_grad(const GridDev& g, const Stencil32<ND>& s, double (&out)[ND]) {
#pragma unroll
    for (int d = 0; d < ND; ++d) out[d] = 0.0;
#pragma unroll
    for (int corner = 0; corner < (1 << ND); ++corner) {
        double weight = 0.0; int lin = 0;
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            const int bit = (corner >> (ND - 1 - d)) & 1;
            const double wd = bit ? s.whi[d] : s.wlo[d];
            weight = (d == 0) ? wd : weight * wd;
            lin += bit ? s.hi[d] : s.lo[d];
        }
        if (ND == 4) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(g.grads) + lin);
            out[0] = out[0] + weight * (double)v.x; out[1] = out[1] + weight * (double)v.y;
            out[2] = out[2] + weight * (double)v.z; out[3] = out[3] + weight * (double)v.w;
        } else {
#pragma unroll
            for (int d = 0; d < ND; ++d) out[d] = out[d] + weight * (double)__ldg(g.grads + lin * ND + d);
        }
    }
}

struct LeanGrad {
    template <int ND>
    __device__ __forceinline__ static void eval(const GridDev& g, const double (&rel)[ND], double (&out)[ND]) {
        Stencil32<ND> st;
        stencil32_setup<ND>(g, rel, st);
        stencil32_grad<ND>(g, st, out);
    }
};

template <int DYN, class ES>
__device__ __forceinline__ double pair_value(const GridDev& vg, double sep, const ES& s, int i, int j) {
    // safety_filter.py:192-201, 345-354 (+ the value shift of HjDataHandle.update_separation_distance)
    constexpr int ND = DYN == LSM_DYN_DOUBLE_INTEGRATOR ? 4 : 5;
    double rel[ND];
    const double2 pi = s.pos[i], pj = s.pos[j];
    relative_state<DYN>(pi.x, pi.y, s.as2[i], s.as3[i], pj.x, pj.y, s.as2[j], s.as3[j], rel);
    Stencil32<ND> st;
    stencil32_setup<ND>(vg, rel, st);
    if (!st.valid) return INFINITY;
    const double v = stencil32_value<ND>(vg, st);
    if (isnan(v)) return INFINITY;
    return v - (sep - vg.separation_distance);
}

// (a) of the safety filter, pair-parallel over the warp's environments: squared distance and HJ value of every
// (ego, other) pair. A separate function so that the lookup gets its own register allocation (all 2^d
// loads of a stencil in flight) instead of competing with the per-agent state carried by the kernel body.
// `vg` points to the block's shared-memory copy of the grid descriptor.
template <int DYN, int N, int L>
__device__ __noinline__ void pair_phase(const GridDev* __restrict__ vg, EnvShared<DYN, N, L>* Sw, int nenv, int lane) {
    using ES = EnvShared<DYN, N, L>;
    const GridDev g = *vg;
    for (int t = lane; t < nenv * N * N; t += 32) {
        const int el = t / (N * N), r = t - el * (N * N);
        const int i = r / N, j = r - i * N;
        ES& T = Sw[el];
        if (!T.cur_filter || i == j || T.done[0][i] || T.done[0][j]) continue;
        const double2 pi = T.pos[i], pj = T.pos[j];
        const double ddx = pj.x - pi.x, ddy = pj.y - pi.y;
        T.d2aa[r] = ddx * ddx + ddy * ddy;
        T.fval[r] = pair_value<DYN>(g, T.cur_sep, T, i, j);
    }
}

template <int DYN, class ES>
__device__ __forceinline__ void emit_obs_row(const ES& S, int ai, int g /* landmark index */, double x, double y,
                                             double s2, double s3, float* o, int N) {
    // navigation_graph_safe.py:855-875, utils.py:114-137
    const double2 gp = S.pos[N + g];
    if (DYN == LSM_DYN_DOUBLE_INTEGRATOR) {
        o[0] = (float)s2; o[1] = (float)s3; o[2] = (float)(gp.x - x); o[3] = (float)(gp.y - y);
        o[4] = (float)S.lsin[g]; o[5] = (float)S.lcos[g]; o[6] = (float)S.lsp[g];
    } else {
        double rx, ry; rotate_into(gp.x - x, gp.y - y, S.cth[ai], S.sth[ai], rx, ry);
        const double rh = S.lh[g] - s2;
        o[0] = (float)s3; o[1] = (float)rx; o[2] = (float)ry;
        o[3] = (float)sin(rh); o[4] = (float)cos(rh); o[5] = (float)S.lsp[g];
    }
}

// ---------------------------------------------------------------------------------------------
// TMA bulk store (UBLKCP): shared -> global, asynchronous, issued by one lane; frees the warp from
// the store loop and its LSU back-pressure. Source and destination 16-byte aligned, size % 16 == 0.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void bulk_store_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, unsigned bytes) {
    const unsigned saddr = (unsigned)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(gdst), "r"(saddr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int PENDING>
__device__ __forceinline__ void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(PENDING) : "memory"); }

// Graph observation of ONE environment (navigation_graph_safe.py:932-994 + utils.py:139-255), all 32 lanes.
template <int DYN, int N, int L>
__device__ __noinline__ void emit_graph(float* __restrict__ node_obs, float* __restrict__ adj,
                                       const uint32_t* __restrict__ sel_tab, const double r2_lt,
                                       EnvShared<DYN, N, L>& T, float* __restrict__ stage, int ee, int lane, int debug) {
    // arguments by value: a noinline callee would otherwise re-read the kernel parameter block through generic loads
    using ES = EnvShared<DYN, N, L>;
    constexpr int M = ES::M, E = ES::E, W = ES::W, EE = E * E;
    constexpr int F = DYN == LSM_DYN_DOUBLE_INTEGRATOR ? 10 : 11;
    // (a) thresholded distance matrix: d2 in float64 against the exact squared radius; the stored
    //     float32 value is d2f * rsqrt(d2f). One lane-task per unordered entity pair, enumerated without a
    //     table as (a, a + d mod E) for d = 1 .. E/2 (for even E the last distance only needs a < E/2).
    for (int e = lane; e < E; e += 32) T.dthr[e * E + e] = 0.0f;
    constexpr int NPAIR = E * (E - 1) / 2;
    auto pair_task = [&](int p) {
        const int dm1 = p / E, a = p - dm1 * E;
        int b = a + dm1 + 1; if (b >= E) b -= E;
        const double2 pa = T.pos[a], pb = T.pos[b];
        const double dx = pa.x - pb.x, dy = pa.y - pb.y;
        const double d2 = dx * dx + dy * dy;
        const float d2f = fmaxf((float)d2, 1.0e-30f);
        float rs;
        asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(d2f));
        const float v = (d2 < r2_lt && d2 > 0.0) ? d2f * rs : 0.0f;
        T.dthr[a * E + b] = v; T.dthr[b * E + a] = v;
    };
    {
        int p = lane;
        for (; p + 32 < NPAIR; p += 64) { pair_task(p); pair_task(p + 32); }   // two independent tasks in flight
        if (p < NPAIR) pair_task(p);
    }
    // (b) disconnected-entity bit masks before / after this step's goal updates (ballots)
    unsigned any_change = 0u, any_disc = 0u;
#pragma unroll
    for (int w = 0; w < W; ++w) {
        const int e = w * 32 + lane;
        bool dpre = false, dpost = false;
        if (e < N) { dpre = T.done[0][e] != 0; dpost = T.done[1][e] != 0; }
        else if (e < E) {
            const int m = e - N, order = m / N, owner = m - order * N;
            dpre = T.reached[0][owner] > order; dpost = T.reached[1][owner] > order;
        }
        const unsigned bpre = __ballot_sync(0xffffffffu, dpre), bpost = __ballot_sync(0xffffffffu, dpost);
        if (lane == 0) { T.disc[0][w] = bpre; T.disc[1][w] = bpost; }
        any_change |= (bpre ^ bpost); any_disc |= bpost;
    }
    __syncwarp();
    if (any_disc != 0u) {
        for (int k = lane; k < N * W; k += 32) {
            const int w = k % W;
            const unsigned sel = sel_tab[k];
            T.keepm[k] = ~((T.disc[1][w] & sel) | (T.disc[0][w] & ~sel));
        }
    }
    __syncwarp();
    // (d) adjacency (issued BEFORE the node rows so that the asynchronous copies overlap their computation)
    if (!(debug & 8)) {
        float* abase = adj + (size_t)ee * (N * EE);
        if (E % 2 == 0 && any_change == 0u && (E % 4 != 0 || (debug & 16))) {
            // every observer sees the same matrix: mask it once in place, then N bulk copies. Used when rows are not
            // 16-byte multiples (E % 4 != 0, e.g. cfg3's E = 30), where it beats the scalar store loop; with E % 4 == 0
            // the float4 loop below is faster (N back-to-back UBLKCP issues stall ~400 cycles each).
            if (any_disc != 0u) {
                for (int idx = lane; idx < EE; idx += 32) {
                    const int a = idx / E, b2 = idx - a * E;
                    const bool keep = ((T.keepm[a >> 5] >> (a & 31)) & 1u) && ((T.keepm[b2 >> 5] >> (b2 & 31)) & 1u);
                    if (!keep) T.dthr[idx] = 0.0f;
                }
            }
            bulk_store_fence();
            __syncwarp();
            if (lane < N) { bulk_store(abase + lane * EE, T.dthr, (unsigned)EE * 4u); bulk_store_commit(); }
        } else if (E % 4 == 0) {
            constexpr int CPR = E / 4, CHUNKS = EE / 4;
            for (int ch = lane; ch < CHUNKS; ch += 32) {
                const int a = ch / CPR, b4 = (ch - a * CPR) * 4;
                const float4 v = *reinterpret_cast<const float4*>(T.dthr + ch * 4);
                float* dst = abase + ch * 4;
                if (any_disc == 0u) {               // nothing disconnected: the same chunk for every observer
#pragma unroll
                    for (int i = 0; i < N; ++i) __stcs(reinterpret_cast<float4*>(dst + i * EE), v);
                } else if (any_change == 0u) {      // no goal update this step: one mask for every observer
                    const bool ka = (T.keepm[a >> 5] >> (a & 31)) & 1u;
                    const unsigned nib = ka ? ((T.keepm[b4 >> 5] >> (b4 & 31)) & 0xFu) : 0u;
                    float4 o;
                    o.x = (nib & 1u) ? v.x : 0.0f; o.y = (nib & 2u) ? v.y : 0.0f;
                    o.z = (nib & 4u) ? v.z : 0.0f; o.w = (nib & 8u) ? v.w : 0.0f;
#pragma unroll
                    for (int i = 0; i < N; ++i) __stcs(reinterpret_cast<float4*>(dst + i * EE), o);
                } else {
#pragma unroll
                    for (int i = 0; i < N; ++i) {
                        const bool ka = (T.keepm[i * W + (a >> 5)] >> (a & 31)) & 1u;
                        const unsigned nib = ka ? ((T.keepm[i * W + (b4 >> 5)] >> (b4 & 31)) & 0xFu) : 0u;
                        float4 o;
                        o.x = (nib & 1u) ? v.x : 0.0f; o.y = (nib & 2u) ? v.y : 0.0f;
                        o.z = (nib & 4u) ? v.z : 0.0f; o.w = (nib & 8u) ? v.w : 0.0f;
                        __stcs(reinterpret_cast<float4*>(dst + i * EE), o);
                    }
                }
            }
        } else {
            for (int idx = lane; idx < EE; idx += 32) {
                const int a = idx / E, b2 = idx - a * E;
                const float v = T.dthr[idx];
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    const bool keep = any_disc == 0u ||
                                      (((T.keepm[i * W + (a >> 5)] >> (a & 31)) & 1u) &&
                                       ((T.keepm[i * W + (b2 >> 5)] >> (b2 & 31)) & 1u));
                    __stcs(abase + i * EE + idx, keep ? v : 0.0f);
                }
            }
        }
    }
    // (c) node features: one lane per (observer, entity) row. Rows are 40 / 44 bytes, so they are staged in a
    //     per-warp double-buffered shared-memory buffer (32 rows at a time) and flushed with fully coalesced vector stores.
    if (!(debug & 4)) {
        float* nbase = node_obs + (size_t)ee * (N * E * F);
        constexpr int ROWS = N * E, CH = 32;
        constexpr int VEC = ((ROWS * F) % 4 == 0 && (CH * F) % 4 == 0) ? 4 : (((ROWS * F) % 2 == 0 && (CH * F) % 2 == 0) ? 2 : 1);
        auto node_row = [&](int r, float* o) {
            const int i = r / E, e = r - i * E;
            const double2 pi = T.pos[i], vi = T.vel[N + i];
            const bool is_agent = e < N;
            const int sel = (e <= i) ? N : 0;     // agents <= i are seen after their own update
            const int vidx = is_agent ? sel + e : 2 * N;
            const int gidx = is_agent ? E + sel + e : e;
            const int cidx = is_agent ? M + sel + e : e - N;
            if (DYN == LSM_DYN_DOUBLE_INTEGRATOR) {
                // utils.py:201-255: [p_e - p_i, v_e - v_i, goal_e - p_i, sin gh, cos gh, gspeed, type]
                const double2 pe = T.pos[e], ve = T.vel[vidx], ge = T.pos[gidx];
                const float4 cc = T.cst[cidx];
                float2* o2 = reinterpret_cast<float2*>(o);   // rows are 40 B: 8 B aligned
                o2[0] = make_float2((float)(pe.x - pi.x), (float)(pe.y - pi.y));
                o2[1] = make_float2((float)(ve.x - vi.x), (float)(ve.y - vi.y));
                o2[2] = make_float2((float)(ge.x - pi.x), (float)(ge.y - pi.y));
                o2[3] = make_float2(cc.x, cc.y);
                o2[4] = make_float2(cc.z, cc.w);
            } else {
                const double ci = T.cth[i], si = T.sth[i];
                if (e < N) {
                    const int g = T.goal[sel ? 1 : 0][e];
                    const double2 pe = T.pos[e], ve = T.vel[vidx], ge = T.pos[gidx];
                    double rx, ry, gx, gy;
                    rotate_into(pe.x - pi.x, pe.y - pi.y, ci, si, rx, ry);
                    rotate_into(ge.x - pi.x, ge.y - pi.y, ci, si, gx, gy);
                    const double ce = T.cth[e], se = T.sth[e];
                    o[0] = (float)rx; o[1] = (float)ry; o[2] = (float)norm2(ve.x - vi.x, ve.y - vi.y);
                    o[3] = (float)(se * ci - ce * si); o[4] = (float)(ce * ci + se * si);
                    o[5] = (float)gx; o[6] = (float)gy;
                    o[7] = (float)(T.lsin[g] * ci - T.lcos[g] * si); o[8] = (float)(T.lcos[g] * ci + T.lsin[g] * si);
                    o[9] = (float)T.lsp[g]; o[10] = 0.0f;
                } else {
                    const int m = e - N;
                    const double2 pe = T.pos[e];
                    double rx, ry;
                    rotate_into(pe.x - pi.x, pe.y - pi.y, ci, si, rx, ry);
                    const float sh = (float)(T.lsin[m] * ci - T.lcos[m] * si), ch = (float)(T.lcos[m] * ci + T.lsin[m] * si);
                    o[0] = (float)rx; o[1] = (float)ry; o[2] = (float)T.spd_post[i];
                    o[3] = sh; o[4] = ch; o[5] = (float)rx; o[6] = (float)ry; o[7] = sh; o[8] = ch;
                    o[9] = (float)T.lsp[m]; o[10] = 1.0f;
                }
            }
        };
        constexpr bool BULK = (VEC == 4);
        int chunk = 0;
        for (int r0 = 0; r0 < ROWS; r0 += CH, ++chunk) {
            float* buf = stage + (chunk & 1) * (CH * F);
            if (BULK && chunk >= 2) {       // the copy that last read this buffer must have finished reading it
                if (lane == 0) bulk_store_wait_read<1>();
                __syncwarp();
            }
            const int ra = r0 + lane;
            if (ra < ROWS) node_row(ra, buf + lane * F);
            const int nfl = ((ROWS - r0) < CH ? (ROWS - r0) : CH) * F;   // floats in this chunk
            float* gdst = nbase + r0 * F;
            if (BULK) {
                bulk_store_fence();
                __syncwarp();
                if (lane == 0) { bulk_store(gdst, buf, (unsigned)nfl * 4u); bulk_store_commit(); }
            } else {
                __syncwarp();
                if (VEC == 2) {
                    for (int q = lane; q < nfl / 2; q += 32)
                        __stcs(reinterpret_cast<float2*>(gdst) + q, reinterpret_cast<const float2*>(buf)[q]);
                } else {
                    for (int q = lane; q < nfl; q += 32) __stcs(gdst + q, buf[q]);
                }
                __syncwarp();
            }
        }
    }
    // every bulk copy issued by this call has finished READING shared memory before the records are reused
    bulk_store_wait_read<0>();
    __syncwarp();
}

// Scenario.random_scenario for ONE environment, executed by the env's leader lane
// (navigation_graph_safe.py:1199-1367, utils.py:39-68); Philox stream keyed by (seed, env, reset_count).
template <int DYN, int N, int L>
__device__ __noinline__ void sample_scenario(const KParams& kp, EnvShared<DYN, N, L>& S, int env, int reset_count,
                                            double ratio) {
    using ES = EnvShared<DYN, N, L>;
    constexpr int M = ES::M;
    const lsm_config& c = kp.c;
    const bool use_filter_arg = (c.flags & LSM_FLAG_USE_SAFETY_FILTER) != 0;
    Rng r; r.init(kp.seed, (uint32_t)(kp.b.env_id_base + env), (uint32_t)reset_count);
    const double ws = c.world_size;
    double cra = ratio_sloped(ratio, 0.25, 0.75);
    if (use_filter_arg) cra = 1.0;
    for (int i = 0; i < N; ++i) {
        if (DYN == LSM_DYN_DOUBLE_INTEGRATOR) {
            const double px = r.uniform(-0.8 * ws, 0.8 * ws);
            const double py = r.uniform(-0.8 * ws, 0.8 * ws);
            S.pos[i] = make_double2(px, py);
            S.as2[i] = 0.0; S.as3[i] = 0.0;
        } else {
            const double xmin = -0.5 * ws;
            const double xmax = 0.25 * ws * cra + 0.0 * (1.0 - cra) * ws;
            const double ry = r.uniform(-0.5 * ws, 0.5 * ws);
            const double rx = r.uniform(xmin, xmax);
            S.pos[i] = make_double2(rx, ry);
            const double sp = r.uniform(c.goal_speed_min, c.goal_speed_max);
            S.as2[i] = r.uniform(0.0, 2.0 * kPi);
            S.as3[i] = sp;
        }
    }
    double2* lp = S.pos + N;    // landmark positions, slot l*N + i
    for (int i = 0; i < N; ++i) {
        double xlo, xhi, ylo, yhi, min_d, max_d;
        if (DYN == LSM_DYN_DOUBLE_INTEGRATOR) {
            xlo = -0.5 * ws; xhi = 0.5 * ws; ylo = -0.5 * ws; yhi = 0.5 * ws;
            min_d = 0.25 * c.coordination_range; max_d = 0.75 * c.coordination_range;
        } else {
            const double yw = 0.1 * (1.0 - cra) + 0.5 * cra;
            xlo = 0.0; xhi = 0.75 * ws; ylo = -yw * ws; yhi = yw * ws;
            min_d = 0.5 * c.coordination_range; max_d = c.coordination_range;
        }
        for (int l = 0; l < L; ++l) {
            double gx = 0.0, gy = 0.0;
            if (l > 0) {
                for (int j = 0; j < 1000; ++j) {
                    gx = r.uniform(xlo, xhi); gy = r.uniform(ylo, yhi);
                    double dm = INFINITY;
                    for (int k = 0; k < l; ++k) {
                        const double d = norm2(lp[k * N + i].x - gx, lp[k * N + i].y - gy);
                        if (d < dm) dm = d;
                    }
                    if (dm > min_d && dm < max_d) break;
                }
            } else { gx = r.uniform(xlo, xhi); gy = r.uniform(ylo, yhi); }
            lp[l * N + i] = make_double2(gx, gy);
        }
        if (i > 0) for (int l = 0; l < L; ++l) if (r.uniform(0.0, 1.0) < 0.5) lp[l * N + i] = lp[l * N + i - 1];
        if (DYN != LSM_DYN_DOUBLE_INTEGRATOR) {
            if (lp[i].x > lp[N + i].x) { const double2 t = lp[i]; lp[i] = lp[N + i]; lp[N + i] = t; }
        }
        for (int l = 0; l < L - 1; ++l)
            S.lh[l * N + i] = atan2(lp[(l + 1) * N + i].y - lp[l * N + i].y, lp[(l + 1) * N + i].x - lp[l * N + i].x);
        const double last_heading = S.lh[(L - 2) * N + i];
        const double cr = use_filter_arg ? 1.0 : ratio_sloped(ratio, 0.25, 0.75);
        if (DYN != LSM_DYN_DOUBLE_INTEGRATOR) {
            for (int l = 0; l < L; ++l) S.lsp[l * N + i] = c.goal_speed_max;
        } else {
            for (int l = 0; l < L; ++l) S.lsp[l * N + i] = r.uniform(c.goal_speed_min, c.goal_speed_max);
            const double var = r.uniform(0.0, 1.0);
            if (!(var < pymin(cr, 1.0 - 0.2))) {
                for (int l = 0; l < L; ++l) S.lsp[l * N + i] = c.goal_speed_max;
                S.lsp[(L - 1) * N + i] = c.goal_speed_min;
            }
        }
        for (int l = 0; l < L - 1; ++l) {
            const double pr = (DYN == LSM_DYN_DOUBLE_INTEGRATOR) ? cr * 0.25 * kPi : cra * 0.1 * kPi;
            S.lh[l * N + i] += r.uniform(-pr, pr);
        }
        S.lh[(L - 1) * N + i] = last_heading;
    }
    for (int m = 0; m < M; ++m) {
        const double sv = sin(S.lh[m]), cv = cos(S.lh[m]);
        S.lsin[m] = sv; S.lcos[m] = cv;
        S.cst[m] = make_float4((float)sv, (float)cv, (float)S.lsp[m], 1.0f);
    }
}

template <int DYN, int N, int L, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) lsm_spec_kernel(const __grid_constant__ KParams kp) {
    using ES = EnvShared<DYN, N, L>;
    constexpr int M = ES::M, E = ES::E, W = ES::W, EE = E * E;
    constexpr int G = Pow2<N>::value;
    constexpr int Dobs = DYN == LSM_DYN_DOUBLE_INTEGRATOR ? 7 : 6;
    constexpr int F = DYN == LSM_DYN_DOUBLE_INTEGRATOR ? 10 : 11;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ GridDev s_vg;                      // block copy of the value-grid descriptor for pair_phase
    if (threadIdx.x == 0) s_vg = kp.vg;
    __syncthreads();
    const lsm_config& c = kp.c;
    const int lane = threadIdx.x & 31;
    const int warp_in_block = threadIdx.x >> 5;
    const int warps_per_block = blockDim.x >> 5;
    const int EPW = kp.EPW;                       // launch choice, 1 .. 32/G
    const int n = (int)kp.b.num_envs;
    const int le = lane / G;
    const int ai = lane - le * G;
    ES* const Sw = reinterpret_cast<ES*>(smem_raw) + warp_in_block * EPW;   // this warp's env records
    const bool lane_has_env = le < EPW;
    // per-warp staging buffer for node-feature rows (after all environment records of the block)
    float* const stage = reinterpret_cast<float*>(smem_raw + (size_t)warps_per_block * EPW * sizeof(ES)) + warp_in_block * (2 * 32 * F);
    ES& S = Sw[lane_has_env ? le : 0];
    const unsigned group_mask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((le * G) & 31));
    const bool use_filter_arg = (c.flags & LSM_FLAG_USE_SAFETY_FILTER) != 0;
    const int ngroups = kp.ngroups;               // ceil(n / EPW), from the host
    const size_t fstride = (size_t)n * N;         // elements between two fields of the agent SoA

    for (int grp = blockIdx.x * warps_per_block + warp_in_block; grp < ngroups; grp += gridDim.x * warps_per_block) {
        const int env0 = grp * EPW;
        const int env = env0 + le;
        bool env_on = lane_has_env && env < n;
        if (kp.mode == MODE_RESET && kp.env_mask != nullptr && env_on) env_on = kp.env_mask[env] != 0;
        const bool agent_on = env_on && ai < N;
        if (__ballot_sync(0xffffffffu, env_on) == 0u) continue;
        const int nenv = (n - env0) < EPW ? (n - env0) : EPW;   // envs of this group

        // ---------------- P0: load ----------------
        double x = 0, y = 0, s2 = 0, s3 = 0, p_dist = 0, state_time = 0, min_rel = INFINITY, goal_min_time = INFINITY;
        double times_req = -1, dists_goal = -1, dist_left = -1, ep_travel_dist = 0, ep_min_dist = INFINITY, action_diff = 0;
        int reached = 0, done = 0, safety_filtered = 0, deconflict = -1, ncoll = 0;
        int ep_len = 0, ep_conflict = 0, ep_multi = 0, ep_done = 0;
        int current_step = 0, reset_count = 0, parity = 0;
        double ratio = 0.0;
        double* const af = kp.b.agent_f64 + ((size_t)env * N + ai);
        int* const aip = kp.b.agent_i32 + ((size_t)env * N + ai);
        if (env_on) {
            current_step = kp.b.env_i32[(size_t)LSM_EI_CURRENT_STEP * n + env];
            reset_count = kp.b.env_i32[(size_t)LSM_EI_RESET_COUNT * n + env];
            parity = kp.b.env_i32[(size_t)LSM_EI_PARITY * n + env];
            ratio = kp.b.env_f64[(size_t)LSM_EF_CURRICULUM_RATIO * n + env];
        }
        if (agent_on) {
            x = af[LSM_AF_X * fstride]; y = af[LSM_AF_Y * fstride]; s2 = af[LSM_AF_S2 * fstride]; s3 = af[LSM_AF_S3 * fstride];
            p_dist = af[LSM_AF_P_DIST * fstride]; state_time = af[LSM_AF_STATE_TIME * fstride];
            min_rel = af[LSM_AF_MIN_REL_DIST * fstride]; goal_min_time = af[LSM_AF_GOAL_MIN_TIME * fstride];
            times_req = af[(parity ? LSM_AF_TIMES_REQ_B : LSM_AF_TIMES_REQ_A) * fstride];
            dists_goal = af[(parity ? LSM_AF_DISTS_GOAL_B : LSM_AF_DISTS_GOAL_A) * fstride];
            dist_left = af[LSM_AF_DIST_LEFT * fstride]; ep_travel_dist = af[LSM_AF_EP_TRAVEL_DIST * fstride];
            ep_min_dist = af[LSM_AF_EP_MIN_DIST * fstride]; action_diff = af[LSM_AF_ACTION_DIFF * fstride];
            reached = aip[LSM_AI_REACHED * fstride]; done = aip[LSM_AI_DONE * fstride];
            safety_filtered = aip[LSM_AI_SAFETY_FILTERED * fstride]; deconflict = aip[LSM_AI_DECONFLICT_IDX * fstride];
            ncoll = aip[LSM_AI_NUM_COLLISIONS * fstride]; ep_len = aip[LSM_AI_EP_TRAVEL_LEN * fstride];
            ep_conflict = aip[LSM_AI_EP_CONFLICT * fstride]; ep_multi = aip[LSM_AI_EP_MULTI * fstride];
            ep_done = aip[LSM_AI_EP_DONE * fstride];
        }
        // landmark tables of the group's environments: contiguous runs per field
        {
            const int total = nenv * M;
            const size_t lstride = (size_t)n * M;
            const double* src = kp.b.landmarks + (size_t)env0 * M;
            for (int idx = lane; idx < total; idx += 32) {
                const int el = idx / M, m = idx - el * M;
                ES& T = Sw[el];
                const double lxv = src[LSM_LF_X * lstride + idx], lyv = src[LSM_LF_Y * lstride + idx];
                const double lhv = src[LSM_LF_HEADING * lstride + idx], lsv = src[LSM_LF_SPEED * lstride + idx];
                const double sv = src[LSM_LF_SIN * lstride + idx], cv = src[LSM_LF_COS * lstride + idx];
                T.pos[N + m] = make_double2(lxv, lyv);
                T.lh[m] = lhv; T.lsp[m] = lsv; T.lsin[m] = sv; T.lcos[m] = cv;
                T.cst[m] = make_float4((float)sv, (float)cv, (float)lsv, 1.0f);   // landmark rows: type 1
            }
        }
        Curriculum q = curriculum(kp, ratio);
        const int lvl = (int)(q.stair * 4.0);     // curriculum stair level: index of the squared-threshold tables
        if (agent_on) {
            S.pos[ai] = make_double2(x, y); S.as2[ai] = s2; S.as3[ai] = s3;
            S.done[0][ai] = done; S.reached[0][ai] = reached;
            if (ai == 0) { S.cur_sep = q.sep; S.cur_filter = q.world_filter ? 1 : 0; S.vel[2 * N] = make_double2(0.0, 0.0); }
        }
        __syncwarp();

        bool all_done_env = false;
        const unsigned any_filter = __ballot_sync(0xffffffffu, env_on && q.world_filter);

        if (kp.mode == MODE_STEP) {
            // ---------------- P1: action decode, safety filter, dynamics ----------------
            current_step += 1;
            double raw0 = 0.0, raw1 = 0.0;
            if (agent_on) {
                int idx;
                if (kp.action_idx != nullptr) idx = kp.action_idx[(size_t)env * N + ai];
                else {   // np.argmax over the one-hot row: first maximum
                    const float* row = kp.action_onehot + ((size_t)env * N + ai) * LSM_NUM_ACTIONS;
                    idx = 0; float best = row[0];
                    for (int k = 1; k < LSM_NUM_ACTIONS; ++k) { const float v = row[k]; if (v > best) { best = v; idx = k; } }
                }
                const int i0 = idx / 5, i1 = idx - i0 * 5;
                raw0 = c.act_tab0[i0]; raw1 = c.act_tab1[i1];
                S.rawx[ai] = raw0; S.rawy[ai] = raw1;
            }
            __syncwarp();
            double safe0 = raw0, safe1 = raw1;
            for (int it = 0; it < c.num_internal_step; ++it) {
                if (any_filter != 0u && !(kp.debug & 2)) {
                    pair_phase<DYN, N, L>(&s_vg, Sw, nenv, lane);
                    __syncwarp();
                }
                if (agent_on && q.world_filter) {
                    // (b) np.argmin over the others (first minimum, ascending agent index), then resolve
                    int filt = 0, dec = -1;
                    safe0 = raw0; safe1 = raw1;
                    if (!done) {
                        double best_d2 = 0.0, best_v = 0.0; int kd = -1, kv = -1;
#pragma unroll
                        for (int j = 0; j < N; ++j) {
                            if (j == ai || S.done[0][j]) continue;
                            const double d2 = S.d2aa[ai * N + j], v = S.fval[ai * N + j];
                            if (kd < 0 || d2 < best_d2) { kd = j; best_d2 = d2; }
                            if (kv < 0 || v < best_v) { kv = j; best_v = v; }
                        }
                        if (kv >= 0) {
                            dec = kv;
                            // `min distance > coordination_range` as an exact test on the squared distance
                            const double best_d = (best_d2 >= kp.r2_gt) ? INFINITY : 0.0;
                            const double2 po = S.pos[kv];
                            filter_resolve<DYN, LeanGrad>(kp, best_d, best_v, !isinf(best_v), x, y, s2, s3, po.x, po.y, S.as2[kv], S.as3[kv],
                                                raw0, raw1, S.rawx[kv], S.rawy[kv], safe0, safe1, filt);
                        }
                    }
                    deconflict = dec; safety_filtered = filt;
                }
                __syncwarp();   // everyone has read the pre-integration states
                if (agent_on) {
                    const double d0 = raw0 - safe0, d1 = raw1 - safe1;
                    action_diff = sqrt(d0 * d0 + d1 * d1);
                    if (!done) integrate<DYN>(x, y, s2, s3, safe0, safe1, c.dt, p_dist, state_time);
                    S.pos[ai] = make_double2(x, y); S.as2[ai] = s2; S.as3[ai] = s3;
                }
                __syncwarp();
            }
            // ---------------- P2: squared agent-agent distances (pair-parallel), goal / reward / done ----------------
            for (int t = lane; t < nenv * N * N; t += 32) {
                const int el = t / (N * N), r = t - el * (N * N);
                const int i = r / N, j = r - i * N;
                ES& T = Sw[el];
                const double2 pi = T.pos[i], pj = T.pos[j];
                const double dx = pi.x - pj.x, dy = pi.y - pj.y;
                T.d2aa[r] = dx * dx + dy * dy;
            }
            __syncwarp();
            int goal_pre = 0, goal_post = 0, reached_post = reached, done_post = done;
            double rew = 0.0;
            double vpx = 0, vpy = 0, vqx = 0, vqy = 0;
            double theta = 0, speed = 0;
            bool reached_now = false;
            if (agent_on) {
                theta = theta_of<DYN>(s2, s3); speed = speed_of<DYN>(s2, s3);
                if (DYN == LSM_DYN_DOUBLE_INTEGRATOR) { vpx = s2; vpy = s3; }
                else { const double ct = cos(s2), st = sin(s2); vpx = s3 * ct; vpy = s3 * st; S.cth[ai] = ct; S.sth[ai] = st; }
                goal_pre = goal_index(reached, ai, N, M);
                const double2 gp = S.pos[N + goal_pre];
                const double gx = gp.x, gy = gp.y, gh = S.lh[goal_pre], gs = S.lsp[goal_pre];
                emit_obs_row<DYN>(S, ai, goal_pre, x, y, s2, s3, kp.b.obs + ((size_t)env * N + ai) * Dobs, N);
                // reward_reach_goal: navigation_graph_safe.py:691-791
                const double he = direction_alignment_error(theta, gh);
                const double hpr = 1.0 - clipd(he / q.heading_thresh, 0.0, 1.0);
                const double se = fabs(speed - gs);
                const double sen = clipd(se / q.speed_thresh, 0.0, 1.0);
                double cra = ratio_sloped(ratio, 0.25, 0.75);
                if (use_filter_arg) cra = 1.0;
                reached_now = goal_reached<DYN>(x, y, theta, speed, gx, gy, gh, gs, q);
                if (reached_now) {
                    const double spr = 1.0 - sen;
                    const double pdx = gx - x, pdy = gy - y;
                    double cte = pdx * sin(theta) - pdy * cos(theta);
                    const double nrm = norm2(pdx, pdy);
                    cte = fabs(cte) / (nrm > 1e-6 ? nrm : 1e-6);
                    cte = clipd(cte, 0.0, 1.0);
                    const double pr = hpr * spr * (1.0 - cte);
                    double goal_rew;
                    if (DYN == LSM_DYN_DOUBLE_INTEGRATOR) goal_rew = c.goal_rew * pr;
                    else goal_rew = c.goal_rew * (pr * cra + (1.0 - cra));
                    if (!done) rew += goal_rew;
                }
                if (!done) {
                    if (DYN == LSM_DYN_DOUBLE_INTEGRATOR) {
                        if (!use_filter_arg) {   // utils.py:323-349
                            const double cg = cos(gh), sg = sin(gh);
                            double rpx, rpy, rvx, rvy;
                            rotate_into(x - gx, y - gy, cg, sg, rpx, rpy);
                            const double dist = norm2(rpx, rpy);
                            const double ang = atan2(rpy, rpx);
                            const double ang_range = kPi / 6;
                            rotate_into(s2 - 0.0, s3 - 0.0, cg, sg, rvx, rvy);
                            const double rh = magnetic_heading(rpx, rpy, 2.0 * q.dist_thresh);
                            double ref_speed = pymax(gs, 0.1);
                            const double dr = clipd(dist / 1.5, 0.0, 1.0);
                            ref_speed = ref_speed * (1.0 - dr) + 1.0 * dr;
                            const double ex = rvx - ref_speed * cos(rh), ey = rvy - ref_speed * sin(rh);
                            const double err = norm2(ex, ey);
                            double pen;
                            if (cos(ang) < cos(ang_range)) pen = err;
                            else {
                                const double ar = clipd((cos(ang) - cos(ang_range)) / (1.0 - cos(ang_range)), 0.0, 1.0);
                                pen = err * (1.0 - ar) + dist * ar;
                            }
                            double hap = 3.0 * pen;
                            hap = clipd(1.0 - q.sloped, 0.0, 1.0) * hap;
                            rew -= hap;
                        }
                        if (use_filter_arg) rew -= 1.0; else rew -= 1.0 * q.sloped;
                    } else {
                        double rpx, rpy;
                        rotate_into(x - gx, y - gy, cos(gh), sin(gh), rpx, rpy);
                        const double rs[4] = { rpx, rpy, theta - gh, speed };
                        Stencil32<4> st;
                        stencil32_setup<4>(kp.tg, rs, st);
                        double ttr = st.valid ? stencil32_value<4>(kp.tg, st) : NAN;
                        if (isnan(ttr)) ttr = kp.tg.ttr_max;
                        rew -= 0.04 * ttr;
                        rew -= sen * cra;
                    }
                }
                // update_reached_goal_and_done (+ freeze_agent): navigation_graph_safe.py:658-675, 1091-1099
                if (reached_now && !done) reached_post = reached + 1;
                done_post = done;
                vqx = vpx; vqy = vpy;
                double s2q = s2, s3q = s3;
                if (reached_post >= L) {
                    done_post = 1;
                    if (DYN == LSM_DYN_DOUBLE_INTEGRATOR) { s2q = 0.0; s3q = 0.0; vqx = 0.0; vqy = 0.0; }
                    else { s3q = 0.0; vqx = s3q * S.cth[ai]; vqy = s3q * S.sth[ai]; }
                }
                goal_post = goal_index(reached_post, ai, N, M);
                S.vel[ai] = make_double2(vpx, vpy); S.vel[N + ai] = make_double2(vqx, vqy);
                S.pos[E + ai] = gp; S.pos[E + N + ai] = S.pos[N + goal_post];
                { float4 t = S.cst[goal_pre]; t.w = 0.0f; S.cst[M + ai] = t; }
                { float4 t = S.cst[goal_post]; t.w = 0.0f; S.cst[M + N + ai] = t; }
                S.spd_post[ai] = (DYN == LSM_DYN_DOUBLE_INTEGRATOR) ? 0.0 : s3q;
                S.goal[0][ai] = goal_pre; S.goal[1][ai] = goal_post;
                S.reached[1][ai] = reached_post; S.done[1][ai] = done_post;
                // as2/as3 keep the PRE-update state (the HJ_VALUE term of later agents only reads agents that are
                // not done, whose pre and post states coincide); the lane's registers take the post state
                s2 = s2q; s3 = s3q;
            }
            __syncwarp();
            if (agent_on) {
                // one pass over the other agents: min distance, collisions, episode statistics, proximity rewards.
                // Agent a is seen after its own update if a < i (rewards) / a <= i (statistics), else before.
                double mind2 = INFINITY;                       // core.py:696-709
                double stat_mind2 = INFINITY; int cnt = 0;      // environment.py:1004-1022
                double r_sv = 0.0;                              // navigation_graph_safe.py:793-798
                int pc_count = 0; double pc_pen = 0.0;          // navigation_graph_safe.py:800-823
                const bool want_sv = (c.flags & LSM_FLAG_SAFETY_VIOLATION) != 0;
                const bool want_pc = (c.flags & LSM_FLAG_POTENTIAL_CONFLICT) != 0;
#pragma unroll
                for (int a = 0; a < N; ++a) {
                    if (a == ai) continue;
                    const double d2 = S.d2aa[ai * N + a];
                    const int apre = S.done[0][a], apost = S.done[1][a];
                    if (!done && !apre && d2 < mind2) mind2 = d2;
                    if (d2 < kp.col2_lt) ncoll += 1;           // navigation_graph_safe.py:405-413, 497-501
                    const int adone_r = a < ai ? apost : apre;
                    if (want_sv && d2 < kp.sep2_lt[lvl] && !adone_r) r_sv += q.conflict_rew;
                    if (want_pc && d2 < kp.eng2_lt[lvl] && !adone_r) {
                        const double rd = sqrt(d2);
                        const double2 pa = S.pos[a];
                        const double rx = pa.x - x, ry = pa.y - y;
                        const double closeness = 1.0 - clipd((rd - q.sep) / (q.eng - q.sep), 0.0, 1.0);
                        const double dir = atan2(ry, rx);
                        const double2 va = S.vel[(a < ai ? N : 0) + a];
                        double change = cos(dir) * (va.x - vpx) + sin(dir) * (va.y - vpy);
                        change = fabs(pymin(0.0, change));
                        pc_pen += change * closeness;
                        pc_count += 1;
                    }
                    const int adone_s = a <= ai ? apost : apre;
                    if (!adone_s && d2 < kp.r2_lt && d2 > 0.0) {
                        if (d2 < kp.engref2_lt) cnt++;
                        if (d2 < stat_mind2) stat_mind2 = d2;
                    }
                }
                min_rel = sqrt(mind2);
                if (want_sv) rew += r_sv;
                if (want_pc && pc_count > 1) rew += q.multi_rew * pc_pen;
                if ((c.flags & LSM_FLAG_DIFF_FROM_FILTERED_ACTION) && use_filter_arg) {   // :825-828
                    if (!done) rew += q.diff_rew * action_diff;
                }
                if (c.flags & LSM_FLAG_HJ_VALUE) {                 // :830-837, core.py:459-468
                    double r = 0.0;
                    for (int a = 0; a < N; ++a) {
                        if (a == ai) continue;
                        const int adone = a < ai ? S.done[1][a] : S.done[0][a];
                        if (adone) continue;
                        const double v = pair_value<DYN>(kp.vg, q.sep, S, ai, a);   // as2/as3 hold the pre-update states
                        const double cvp = fabs(pymin(v - 0.4, 0.0));
                        r += q.cvalue_rew * cvp;
                    }
                    rew += r;
                }
                rew = clipd(rew, c.min_reward, c.max_reward);
                if (!done_post) {
                    ep_len += 1;
                    ep_travel_dist += norm2(vqx, vqy) * c.dt;
                    if (stat_mind2 < INFINITY) {
                        if (cnt > 1) ep_multi += 1;
                        if (stat_mind2 < kp.septgt2_lt) ep_conflict += 1;
                        const double mn = sqrt(stat_mind2);
                        if (mn < ep_min_dist) ep_min_dist = mn;
                    }
                }
                if (done_post) ep_done = 1;
                // info_callback state: navigation_graph_safe.py:386-413 (post-update goal and velocity)
                if (times_req == -1.0) {
                    // once times_required is set all three fields are frozen, so the goal test is only needed here
                    const double2 gq = S.pos[N + goal_post];
                    const double dx = x - gq.x, dy = y - gq.y;
                    // same goal and same (unfrozen) state as before the update -> same answer as `reached_now`
                    bool r2 = reached_now;
                    if (goal_post != goal_pre || done_post != done) {
                        const double th2 = theta_of<DYN>(s2, s3), sp2 = speed_of<DYN>(s2, s3);
                        r2 = goal_reached<DYN>(x, y, th2, sp2, gq.x, gq.y, S.lh[goal_post], S.lsp[goal_post], q);
                    }
                    if (r2) times_req = (double)current_step * c.dt;
                    dists_goal = p_dist; dist_left = sqrt(dx * dx + dy * dy);
                }
            }
            if (agent_on && kp.b.reward_individual != nullptr) kp.b.reward_individual[(size_t)env * N + ai] = (float)rew;
            if (c.flags & LSM_FLAG_SHARED_REWARD) {     // environment.py:1032-1037, sequential sum in agent order
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < N; ++k) s += __shfl_sync(0xffffffffu, rew, (le * G + k) & 31);
                rew = s;
            }
            const bool done_out = done_post || (current_step >= c.episode_length);   // environment.py:260-268
            const unsigned not_done = __ballot_sync(0xffffffffu, agent_on && !done_out);
            all_done_env = env_on && ((not_done & group_mask) == 0u);
            if (agent_on) {
                kp.b.reward[(size_t)env * N + ai] = (float)rew;
                kp.b.done[(size_t)env * N + ai] = (uint8_t)done_out;
                reinterpret_cast<double2*>(kp.b.safe_action)[(size_t)env * N + ai] = make_double2(safe0, safe1);
            }
            reached = reached_post; done = done_post;
            parity ^= 1;
        } else {
            if (agent_on) {
                double vx, vy;
                if (DYN == LSM_DYN_DOUBLE_INTEGRATOR) { vx = s2; vy = s3; }
                else { const double ct = cos(s2), st = sin(s2); vx = s3 * ct; vy = s3 * st; S.cth[ai] = ct; S.sth[ai] = st; }
                const int g = goal_index(reached, ai, N, M);
                S.vel[ai] = make_double2(vx, vy); S.vel[N + ai] = make_double2(vx, vy);
                const double2 gp = S.pos[N + g];
                S.pos[E + ai] = gp; S.pos[E + N + ai] = gp;
                float4 cc = S.cst[g]; cc.w = 0.0f;
                S.cst[M + ai] = cc; S.cst[M + N + ai] = cc;
                S.spd_post[ai] = (DYN == LSM_DYN_DOUBLE_INTEGRATOR) ? 0.0 : s3;
                S.goal[0][ai] = g; S.goal[1][ai] = g;
                S.reached[1][ai] = reached; S.done[1][ai] = done;
            }
            __syncwarp();
        }

        // ---------------- P3: reset (graphworker auto-reset or explicit) ----------------
        const bool do_reset = (kp.mode == MODE_STEP && kp.flag && all_done_env) || (kp.mode == MODE_RESET && env_on);
        const bool sample = (kp.mode == MODE_STEP) ? true : (kp.flag != 0);
        const unsigned reset_lanes = __ballot_sync(0xffffffffu, do_reset);
        if (reset_lanes != 0u) {
            double s_len = 0, s_dist = 0, s_done = 0, s_reached = 0, s_conf = 0, s_min = 0, s_multi = 0, mn = INFINITY;
            const double len_i = ep_len == 0 ? 1.0 : (double)ep_len;
            for (int k = 0; k < N; ++k) {
                const int src = (le * G + k) & 31;
                s_len += (double)__shfl_sync(0xffffffffu, ep_len, src);
                s_dist += __shfl_sync(0xffffffffu, ep_travel_dist, src);
                s_done += (double)__shfl_sync(0xffffffffu, ep_done, src);
                s_reached += (double)__shfl_sync(0xffffffffu, reached, src);
                s_conf += __shfl_sync(0xffffffffu, (double)ep_conflict / len_i, src);
                s_multi += __shfl_sync(0xffffffffu, (double)ep_multi / len_i, src);
                const double md = __shfl_sync(0xffffffffu, ep_min_dist, src);
                s_min += md;
                if (md < mn) mn = md;
            }
            if (do_reset && ai == 0) {
                double* out = kp.b.ep_info + (size_t)env * LSM_EP_COUNT;
                out[LSM_EP_TRAVEL_TIME_MEAN] = c.dt * (s_len / N);
                out[LSM_EP_TRAVEL_DISTANCE_MEAN] = s_dist / N;
                out[LSM_EP_DONE_PERCENTAGE] = s_done / N;
                out[LSM_EP_NUM_REACHED_GOAL_MEAN] = s_reached / N;
                out[LSM_EP_CONFLICT_PERCENTAGE] = s_conf / N;
                const double mm = s_min / N;
                out[LSM_EP_MIN_DISTANCE_MEAN] = isinf(mm) ? c.coordination_range : mm;
                out[LSM_EP_MIN_DISTANCE_MIN] = isinf(mn) ? c.coordination_range : mn;
                out[LSM_EP_MULTIPLE_ENGAGEMENT_PERCENTAGE] = s_multi / N;
            }
            if (do_reset) {
                current_step = 0;
                ratio = clipd((double)kp.episode / (double)c.num_total_episode, 0.0, 1.0);
                q = curriculum(kp, ratio);
            }
            if (do_reset && sample && ai == 0) sample_scenario<DYN, N, L>(kp, S, env, reset_count, ratio);
            __syncwarp();
            if (do_reset && agent_on) {
                if (sample) { const double2 p = S.pos[ai]; x = p.x; y = p.y; s2 = S.as2[ai]; s3 = S.as3[ai]; }
                done = 0; reached = 0;
                p_dist = 0.0; state_time = 0.0;
                const double2 g0 = S.pos[N + ai];
                goal_min_time = norm2(x - g0.x, y - g0.y) / c.agent_max_speed;   // navigation_graph_safe.py:525-535
                times_req = -1.0; dists_goal = -1.0; dist_left = -1.0; ncoll = 0;
                ep_len = 0; ep_travel_dist = 0.0; ep_done = 0; ep_conflict = 0; ep_multi = 0; ep_min_dist = INFINITY;
                double vx, vy;
                if (DYN == LSM_DYN_DOUBLE_INTEGRATOR) { vx = s2; vy = s3; }
                else { const double ct = cos(s2), st = sin(s2); vx = s3 * ct; vy = s3 * st; S.cth[ai] = ct; S.sth[ai] = st; }
                S.pos[ai] = make_double2(x, y); S.as2[ai] = s2; S.as3[ai] = s3;
                S.vel[ai] = make_double2(vx, vy); S.vel[N + ai] = make_double2(vx, vy);
                S.pos[E + ai] = g0; S.pos[E + N + ai] = g0;
                float4 cc = S.cst[ai]; cc.w = 0.0f;
                S.cst[M + ai] = cc; S.cst[M + N + ai] = cc;
                S.spd_post[ai] = (DYN == LSM_DYN_DOUBLE_INTEGRATOR) ? 0.0 : s3;
                S.goal[0][ai] = ai; S.goal[1][ai] = ai;
                S.reached[0][ai] = 0; S.reached[1][ai] = 0; S.done[0][ai] = 0; S.done[1][ai] = 0;
                emit_obs_row<DYN>(S, ai, ai, x, y, s2, s3, kp.b.obs + ((size_t)env * N + ai) * Dobs, N);
            }
            if (do_reset && sample) reset_count += 1;
            __syncwarp();
            if (sample) {
                const size_t lstride = (size_t)n * M;
                for (int el = 0; el < nenv; ++el) {
                    if (!((reset_lanes >> ((el * G) & 31)) & 1u)) continue;
                    const ES& T = Sw[el];
                    double* dst = kp.b.landmarks + (size_t)(env0 + el) * M;
                    for (int m = lane; m < M; m += 32) {
                        const double2 p = T.pos[N + m];
                        dst[LSM_LF_X * lstride + m] = p.x; dst[LSM_LF_Y * lstride + m] = p.y;
                        dst[LSM_LF_HEADING * lstride + m] = T.lh[m]; dst[LSM_LF_SPEED * lstride + m] = T.lsp[m];
                        dst[LSM_LF_SIN * lstride + m] = T.lsin[m]; dst[LSM_LF_COS * lstride + m] = T.lcos[m];
                    }
                }
            }
        } else if (kp.mode == MODE_OBSERVE && agent_on) {
            emit_obs_row<DYN>(S, ai, S.goal[0][ai], x, y, s2, s3, kp.b.obs + ((size_t)env * N + ai) * Dobs, N);
        }

        // ---------------- state write-back ----------------
        if (kp.mode != MODE_OBSERVE) {
            if (env_on && ai == 0) {
                kp.b.env_i32[(size_t)LSM_EI_CURRENT_STEP * n + env] = current_step;
                kp.b.env_i32[(size_t)LSM_EI_RESET_COUNT * n + env] = reset_count;
                kp.b.env_i32[(size_t)LSM_EI_PARITY * n + env] = parity;
                kp.b.env_i32[(size_t)LSM_EI_JUST_RESET * n + env] = do_reset ? 1 : 0;
                kp.b.env_f64[(size_t)LSM_EF_CURRICULUM_RATIO * n + env] = ratio;
            }
            if (agent_on) {
                af[LSM_AF_X * fstride] = x; af[LSM_AF_Y * fstride] = y; af[LSM_AF_S2 * fstride] = s2; af[LSM_AF_S3 * fstride] = s3;
                af[LSM_AF_P_DIST * fstride] = p_dist; af[LSM_AF_STATE_TIME * fstride] = state_time;
                af[LSM_AF_MIN_REL_DIST * fstride] = min_rel; af[LSM_AF_GOAL_MIN_TIME * fstride] = goal_min_time;
                af[(parity ? LSM_AF_TIMES_REQ_B : LSM_AF_TIMES_REQ_A) * fstride] = times_req;
                af[(parity ? LSM_AF_DISTS_GOAL_B : LSM_AF_DISTS_GOAL_A) * fstride] = dists_goal;
                if (do_reset) {
                    af[(parity ? LSM_AF_TIMES_REQ_A : LSM_AF_TIMES_REQ_B) * fstride] = times_req;
                    af[(parity ? LSM_AF_DISTS_GOAL_A : LSM_AF_DISTS_GOAL_B) * fstride] = dists_goal;
                }
                af[LSM_AF_DIST_LEFT * fstride] = dist_left; af[LSM_AF_EP_TRAVEL_DIST * fstride] = ep_travel_dist;
                af[LSM_AF_EP_MIN_DIST * fstride] = ep_min_dist; af[LSM_AF_ACTION_DIFF * fstride] = action_diff;
                aip[LSM_AI_REACHED * fstride] = reached; aip[LSM_AI_DONE * fstride] = done;
                aip[LSM_AI_SAFETY_FILTERED * fstride] = safety_filtered; aip[LSM_AI_DECONFLICT_IDX * fstride] = deconflict;
                aip[LSM_AI_NUM_COLLISIONS * fstride] = ncoll; aip[LSM_AI_EP_TRAVEL_LEN * fstride] = ep_len;
                aip[LSM_AI_EP_CONFLICT * fstride] = ep_conflict; aip[LSM_AI_EP_MULTI * fstride] = ep_multi;
                aip[LSM_AI_EP_DONE * fstride] = ep_done;
            }
        }
        __syncwarp();

        // ---------------- P4: graph observation ----------------
        for (int el = 0; el < nenv; ++el) {
            const int ee = env0 + el;
            if (kp.mode == MODE_RESET && kp.env_mask != nullptr && kp.env_mask[ee] == 0) continue;
            if (!(kp.debug & 1)) emit_graph<DYN, N, L>(kp.b.node_obs, kp.b.adj, kp.sel_tab, kp.r2_lt, Sw[el], stage, ee, lane, kp.debug);
            __syncwarp();
        }
        __syncwarp();
    }
}

}  // namespace lsm
