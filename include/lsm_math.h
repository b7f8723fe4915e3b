/*
 * lsm_math.h - self-contained float64 sin / cos / sincos / atan2 shared by the CUDA kernels (device code), the
 * host side of liblsm_b200.so (lookup tables, set_state) and the CPU oracle (oracle/lsm_oracle.c).
 *
 * Why: the reference calls numpy's libm (multiagent/core.py:105-131,179-181, safety_filter.py:277-284,
 * navigation_graph_safe.py:606-656, utils.py:79-349). CUDA's libdevice and glibc round differently in the last
 * bit, and a last-bit difference in a heading or a relative position can flip a thresholded (discrete) output.
 * With ONE implementation made only of IEEE-754 +, -, *, / and integer bit operations - compiled without FMA
 * contraction on both sides (nvcc -fmad=false, gcc -ffp-contract=off) - the CUDA path and the oracle agree BIT
 * FOR BIT on every input, so the parity tests need no tolerance for discrete outputs.
 *
 * Algorithms: the classic fdlibm / msun kernels (Sun Microsystems 1993, freely redistributable): argument
 * reduction by Cody-Waite with a three-part pi/2 (exact for |x| < 2^20 * pi/2; headings here stay below ~1e2 rad),
 * degree-13 / degree-14 minimax polynomials on [-pi/4, pi/4], atan by four-interval reduction + degree-11
 * polynomial in x^2 (the reduction is applied to numerator and denominator, so atan2 needs one division). Error < 1 ulp for
 * sin / cos, < 2 ulp for atan2 (checked against the host libm in
 * tests/test_lsm_math.py). Beyond the
 * Cody-Waite range the reduction falls back to a plain (inaccurate but deterministic) floor-based one; NaN / inf
 * give NaN.
 */
#ifndef LSM_MATH_H
#define LSM_MATH_H

#include <stdint.h>
#include <string.h>

/* LSM_MATH_FN: small helpers, inlined. LSM_MATH_API: the four public functions. In device code they are NOT inlined:
 * the per-agent kernel calls them from ~40 sites and is bound by instruction fetch, so one shared copy of each is faster
 * than 40 inlined reductions + polynomials (the arithmetic - and therefore every result bit - is the same either way). */
#if defined(__CUDACC__)
#define LSM_MATH_FN __host__ __device__ __forceinline__
#define LSM_MATH_API __host__ __device__ __noinline__ inline
#define LSM_MATH_HAS_INL 1
#else
#define LSM_MATH_FN static inline
#define LSM_MATH_API static inline
#endif

LSM_MATH_FN int64_t lsm_m_bits(double x) {
#if defined(__CUDA_ARCH__)
    return __double_as_longlong(x);
#else
    int64_t b; memcpy(&b, &x, sizeof(b)); return b;
#endif
}
LSM_MATH_FN double lsm_m_from_bits(int64_t b) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double(b);
#else
    double x; memcpy(&x, &b, sizeof(x)); return x;
#endif
}
LSM_MATH_FN int32_t lsm_m_hi(double x) { return (int32_t)(lsm_m_bits(x) >> 32); }
LSM_MATH_FN double lsm_m_abs(double x) { return lsm_m_from_bits(lsm_m_bits(x) & 0x7fffffffffffffffLL); }

/* sin on [-pi/4, pi/4]; (x, y) = head / tail of the reduced argument */
LSM_MATH_FN double lsm_m_ksin(double x, double y) {
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03, S3 = -1.98412698298579493134e-04,
                 S4 = 2.75573137070700676789e-06, S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    const double z = x * x;
    const double w = z * z;
    const double r = (S2 + z * (S3 + z * S4)) + (z * w) * (S5 + z * S6);
    const double v = z * x;
    return x - ((z * (0.5 * y - v * r) - y) - v * S1);
}

/* cos on [-pi/4, pi/4] */
LSM_MATH_FN double lsm_m_kcos(double x, double y) {
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03, C3 = 2.48015872894767294178e-05,
                 C4 = -2.75573143513906633035e-07, C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    const double z = x * x;
    double w = z * z;
    const double r = z * (C1 + z * (C2 + z * C3)) + (w * w) * (C4 + z * (C5 + z * C6));
    const double hz = 0.5 * z;
    w = 1.0 - hz;
    return w + (((1.0 - w) - hz) + (z * r - x * y));
}

/* Cody-Waite reduction for |x| < 2^20 * pi/2: x = n * pi/2 + (y0 + y1); returns n mod 4.
 * Written for a SIMT machine: straight-line selects instead of branches wherever lanes of one warp would disagree
 * (quadrant, "no reduction needed", interval of atan) - a divergent branch executes both sides one after the other,
 * two independent polynomial chains interleave. Only the never / rarely taken paths (cancellation near a multiple of
 * pi/2, non-finite and out-of-range arguments) are branches. */
LSM_MATH_FN int lsm_m_rem_medium(double x, int32_t ix, double* y0, double* y1) {
    const double invpio2 = 6.36619772367581382433e-01,
                 pio2_1 = 1.57079632673412561417e+00, pio2_1t = 6.07710050650619224932e-11,
                 pio2_2 = 6.07710050630396597660e-11, pio2_2t = 2.02226624879595063154e-21,
                 pio2_3 = 2.02226624871116645580e-21, pio2_3t = 8.47842766036889956997e-32;
    const double big = 6755399441055744.0;                              /* 1.5 * 2^52: round to nearest integer */
    const double fn = (x * invpio2 + big) - big;
    const int n = (int)fn;
    double r = x - fn * pio2_1;
    double w = fn * pio2_1t;
    const int j = ix >> 20;
    double h = r - w;
    int i = j - ((lsm_m_hi(h) >> 20) & 0x7ff);
    if (i > 16) {                                                       /* 2nd round, good to 118 bits */
        double t = r;
        w = fn * pio2_2;
        r = t - w;
        w = fn * pio2_2t - ((t - r) - w);
        h = r - w;
        i = j - ((lsm_m_hi(h) >> 20) & 0x7ff);
        if (i > 49) {                                                   /* 3rd round, 151 bits */
            t = r;
            w = fn * pio2_3;
            r = t - w;
            w = fn * pio2_3t - ((t - r) - w);
            h = r - w;
        }
    }
    *y0 = h; *y1 = (r - h) - w;
    return n & 3;
}

/* x = n * pi/2 + (y0 + y1), |y0 + y1| <= pi/4 (+ a little); returns n mod 4 */
LSM_MATH_FN int lsm_m_rem_pio2(double x, double* y0, double* y1) {
    const int32_t ix = lsm_m_hi(x) & 0x7fffffff;
    double xr = x;
    int32_t ir = ix;
    if (ix >= 0x413921fb) {                                             /* never taken in contract */
        if (ix >= 0x7ff00000) { *y0 = x - x; *y1 = 0.0; return 0; }    /* inf / NaN -> NaN */
        {   /* |x| >= 2^20 * pi/2 ~ 1.6e6 rad: deterministic, not accurate */
            const double twopi = 6.283185307179586;
            const double big = 6755399441055744.0;
            const double q = x / twopi;
            xr = 0.0;
            if (lsm_m_abs(q) < 2251799813685248.0) xr = x - ((q + big) - big) * twopi;
            ir = lsm_m_hi(xr) & 0x7fffffff;
        }
    }
    {
        double h, t;
        const int n = lsm_m_rem_medium(xr, ir, &h, &t);
        const int small = ir <= 0x3fe921fb;                             /* |x| ~<= pi/4: no reduction (select, not a branch) */
        *y0 = small ? xr : h;
        *y1 = small ? 0.0 : t;
        return small ? 0 : n;
    }
}

/* exactly lsm_sin(x) and lsm_cos(x), one argument reduction; both kernels are always evaluated (independent chains) */
LSM_MATH_FN void lsm_sincos_inl(double x, double* s, double* c) {
    double y0, y1;
    const int32_t ix = lsm_m_hi(x) & 0x7fffffff;
    const int n = lsm_m_rem_pio2(x, &y0, &y1);
    const double ks = lsm_m_ksin(y0, y1), kc = lsm_m_kcos(y0, y1);
    const double a = (n & 1) ? kc : ks;                                 /* n: 0 (s, c)  1 (c, -s)  2 (-s, -c)  3 (-c, s) */
    const double b = (n & 1) ? ks : kc;
    double sv = (n & 2) ? -a : a;
    const double cv = ((n + 1) & 2) ? -b : b;
    if (ix < 0x3e500000) sv = x;                                        /* |x| < 2^-26 */
    *s = sv; *c = cv;
}
LSM_MATH_FN double lsm_sin_inl(double x) { double s, c; lsm_sincos_inl(x, &s, &c); return s; }
LSM_MATH_FN double lsm_cos_inl(double x) { double s, c; lsm_sincos_inl(x, &s, &c); return c; }

/* atan(a / b) for finite a >= 0, b > 0 with ONE division: the four-interval reduction of fdlibm's atan applied to the
 * quotient before it is formed,
 *     t < 7/16: a / b      [7/16, 11/16): (2a - b) / (2b + a)      [11/16, 19/16): (a - b) / (a + b)
 *     [19/16, 39/16): (a - 1.5 b) / (b + 1.5 a)      >= 39/16: -b / a
 * then atan(t) = atanhi[id] + atan(reduced) with the degree-11 polynomial in the reduced argument squared. */
LSM_MATH_FN double lsm_m_atan_ratio(double a, double b) {
    const double hi0 = 4.63647609000806093515e-01, hi1 = 7.85398163397448278999e-01, hi2 = 9.82793723247329054082e-01,
                 hi3 = 1.57079632679489655800e+00;
    const double lo0 = 2.26987774529616870924e-17, lo1 = 3.06161699786838301793e-17, lo2 = 1.39033110312309984516e-17,
                 lo3 = 6.12323399573676603587e-17;
    const double a0 = 3.33333333333329318027e-01, a1 = -1.99999999998764832476e-01, a2 = 1.42857142725034663711e-01,
                 a3 = -1.11111104054623557880e-01, a4 = 9.09088713343650656196e-02, a5 = -7.69187620504482999495e-02,
                 a6 = 6.66107313738753120669e-02, a7 = -5.83357013379057348645e-02, a8 = 4.97687799461593236017e-02,
                 a9 = -3.65315727442169155270e-02, a10 = 1.62858201153657823623e-02;
    /* interval of t = a / b by comparing a with multiples of b (the intervals overlap in validity, so the last-bit
       position of a boundary does not matter) */
    const int g0 = a >= 0.4375 * b, g1 = a >= 0.6875 * b, g2 = a >= 1.1875 * b, g3 = a >= 2.4375 * b;
    double num = a, den = b, ahi = 0.0, alo = 0.0;
    if (g0) { num = (a + a) - b; den = (b + b) + a; ahi = hi0; alo = lo0; }
    if (g1) { num = a - b; den = a + b; ahi = hi1; alo = lo1; }
    if (g2) { num = a - 1.5 * b; den = b + 1.5 * a; ahi = hi2; alo = lo2; }
    if (g3) { num = -b; den = a; ahi = hi3; alo = lo3; }
    {
        const double x = num / den;
        const double z = x * x;
        const double w = z * z;
        const double s1 = z * (a0 + w * (a2 + w * (a4 + w * (a6 + w * (a8 + w * a10)))));
        const double s2 = w * (a1 + w * (a3 + w * (a5 + w * (a7 + w * a9))));
        const double p = x * (s1 + s2);
        const double r_small = x - p;
        const double r_big = ahi - ((p - alo) - x);
        return g0 ? r_big : r_small;
    }
}

LSM_MATH_FN double lsm_atan2_inl(double y, double x) {
    const double pi = 3.1415926535897931160e+00, pi_lo = 1.2246467991473531772e-16;
    const double pi_o_2 = 1.5707963267948965580e+00, pi_o_4 = 7.8539816339744827900e-01;
    const int64_t bx = lsm_m_bits(x), by = lsm_m_bits(y);
    const int32_t hx = (int32_t)(bx >> 32), hy = (int32_t)(by >> 32);
    const int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
    const uint32_t lx = (uint32_t)bx, ly = (uint32_t)by;
    int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);                        /* 2 * sign(x) + sign(y) */
    double z;
    /* special operands first (uniformly not taken on real data) */
    if (x != x || y != y) return x + y;
    if ((iy | (int32_t)ly) == 0) {                                      /* y = +-0 */
        switch (m) { case 0: case 1: return y; case 2: return pi; default: return -pi; }
    }
    if ((ix | (int32_t)lx) == 0) return hy < 0 ? -pi_o_2 : pi_o_2;      /* x = +-0 */
    if (ix == 0x7ff00000) {                                             /* x = +-inf */
        if (iy == 0x7ff00000) {
            switch (m) { case 0: return pi_o_4; case 1: return -pi_o_4; case 2: return 3.0 * pi_o_4; default: return -3.0 * pi_o_4; }
        }
        switch (m) { case 0: return 0.0; case 1: return -0.0; case 2: return pi; default: return -pi; }
    }
    if (iy == 0x7ff00000) return hy < 0 ? -pi_o_2 : pi_o_2;             /* y = +-inf */
    {
        const int32_t k = (iy - ix) >> 20;
        if (k > 60) { z = pi_o_2 + 0.5 * pi_lo; m &= 1; }               /* |y / x| > 2^60 */
        else if (k < -60) z = hx < 0 ? 0.0 : lsm_m_abs(y) / lsm_m_abs(x);   /* tiny quotient: atan(t) = t */
        else z = lsm_m_atan_ratio(lsm_m_abs(y), lsm_m_abs(x));
    }
    {
        const double zn = (m & 2) ? pi - (z - pi_lo) : z;               /* x < 0: reflect */
        return (m & 1) ? -zn : zn;                                      /* y < 0: (z - pi_lo) - pi == -(pi - (z - pi_lo)) */
    }
}

/* the public functions: one shared (not inlined) copy each in device code; `*_inl` above are the same arithmetic,
 * force-inlined, for the two or three call sites on the per-agent kernel's critical path */
#if defined(LSM_MATH_AB_LIBDEVICE) && defined(__CUDA_ARCH__)
/* A/B timing build only (never the product, breaks bit-identity with the oracle): CUDA's own libdevice functions */
#define lsm_sin sin
#define lsm_cos cos
#define lsm_sincos sincos
#define lsm_atan2 atan2
#define lsm_sin_inl sin
#define lsm_cos_inl cos
#define lsm_sincos_inl sincos
#define lsm_atan2_inl atan2
#else
LSM_MATH_API double lsm_sin(double x) { return lsm_sin_inl(x); }
LSM_MATH_API double lsm_cos(double x) { return lsm_cos_inl(x); }
LSM_MATH_API void lsm_sincos(double x, double* s, double* c) { lsm_sincos_inl(x, s, c); }
LSM_MATH_API double lsm_atan2(double y, double x) { return lsm_atan2_inl(y, x); }
#endif

#endif /* LSM_MATH_H */
