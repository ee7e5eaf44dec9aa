/*
 * mctsb200_detmath.h -- the numerical contract of libmctsb200.
 *
 * The reference search loop calls three transcendental functions:
 *   log   in the UCB score          (/root/reference/src/vanilla.jl:368, src/dpw.jl:182)
 *   ^     in the widening criteria  (/root/reference/src/dpw.jl:95,121)  -- handled on the host, see
 *                                    mctsb200.h "widening tables"; never evaluated on the device
 *   cos   in MountainCar's dynamics (POMDPModels, not vendored in the reference)
 *
 * Neither Julia's, glibc's nor CUDA's versions of these are correctly rounded, so a near-tie between
 * two UCB scores can flip between implementations.  To make "bit-exact trees" a meaningful, testable
 * statement, the library and its test oracle evaluate log and cos with the functions below, which are
 * built from IEEE-754 binary64 + - * / only (each individually rounded to nearest, no fused
 * multiply-add), so they return identical bits on the host (gcc -ffp-contract=off) and on the device
 * (explicit __dadd_rn/__dmul_rn/... intrinsics, which ptxas never contracts).
 *
 * Algorithms: the classic argument-reduction + minimax-polynomial schemes (Sun fdlibm lineage;
 * coefficients are the published minimax constants).  Accuracy: < 1 ulp for log on (0, inf);
 * cos has < 1 ulp error for |x| <= 2^20 except in the immediate vicinity of its zeros, where the
 * absolute error stays below 2^-70.  tests/test_detmath.py pins both against glibc/libquadmath.
 */
#ifndef MCTSB200_DETMATH_H
#define MCTSB200_DETMATH_H

#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define MCTSB200_HD __host__ __device__ __forceinline__
#else
#define MCTSB200_HD static inline
#endif

/* Individually-rounded binary64 primitives. */
#if defined(__CUDA_ARCH__)
#define DM_ADD(a, b) __dadd_rn((a), (b))
#define DM_SUB(a, b) __dsub_rn((a), (b))
#define DM_MUL(a, b) __dmul_rn((a), (b))
#define DM_DIV(a, b) __ddiv_rn((a), (b))
#define DM_SQRT(a) __dsqrt_rn((a))
#else
#include <math.h>
#define DM_ADD(a, b) ((a) + (b))
#define DM_SUB(a, b) ((a) - (b))
#define DM_MUL(a, b) ((a) * (b))
#define DM_DIV(a, b) ((a) / (b))
#define DM_SQRT(a) sqrt((a))
#endif

MCTSB200_HD uint64_t dm_bits(double x) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(x);
#else
    uint64_t u;
    memcpy(&u, &x, sizeof u);
    return u;
#endif
}

MCTSB200_HD double dm_from_bits(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double x;
    memcpy(&x, &u, sizeof x);
    return x;
#endif
}

/* Natural logarithm. det_log(0) = -inf, det_log(1) = +0 exactly, det_log(x<0) = NaN. */
MCTSB200_HD double det_log(double x) {
    const double LN2_HI = dm_from_bits(0x3fe62e42fee00000ULL); /* ln2, top 33 bits      */
    const double LN2_LO = dm_from_bits(0x3dea39ef35793c76ULL); /* ln2 - LN2_HI          */
    const double L1 = dm_from_bits(0x3fe5555555555593ULL);
    const double L2 = dm_from_bits(0x3fd999999997fa04ULL);
    const double L3 = dm_from_bits(0x3fd2492494229359ULL);
    const double L4 = dm_from_bits(0x3fcc71c51d8e78afULL);
    const double L5 = dm_from_bits(0x3fc7466496cb03deULL);
    const double L6 = dm_from_bits(0x3fc39a09d078c69fULL);
    const double L7 = dm_from_bits(0x3fc2f112df3e5244ULL);

    uint64_t u = dm_bits(x);
    int k = 0;
    if ((u << 1) == 0) return dm_from_bits(0xfff0000000000000ULL);     /* +-0 -> -inf */
    if (u >> 63) return dm_from_bits(0x7ff8000000000000ULL);           /* negative -> NaN */
    if ((u >> 52) == 0x7ff) return x;                                  /* inf, NaN */
    if ((u >> 52) == 0) {                                              /* subnormal: scale by 2^54 */
        x = DM_MUL(x, 18014398509481984.0);
        u = dm_bits(x);
        k = -54;
    }
    k += (int)(u >> 52) - 1023;
    uint64_t mant = u & 0x000fffffffffffffULL;
    /* m in [sqrt(1/2), sqrt(2)): fold the upper part of [1,2) down by one binade */
    if (mant > 0x6a09e667f3bcdULL) {
        u = mant | 0x3fe0000000000000ULL;
        k += 1;
    } else {
        u = mant | 0x3ff0000000000000ULL;
    }
    const double f = DM_SUB(dm_from_bits(u), 1.0);                     /* exact */
    const double dk = (double)k;
    const double s = DM_DIV(f, DM_ADD(2.0, f));
    const double z = DM_MUL(s, s);
    const double w = DM_MUL(z, z);
    const double t1 = DM_MUL(w, DM_ADD(L2, DM_MUL(w, DM_ADD(L4, DM_MUL(w, L6)))));
    const double t2 = DM_MUL(z, DM_ADD(L1, DM_MUL(w, DM_ADD(L3, DM_MUL(w, DM_ADD(L5, DM_MUL(w, L7)))))));
    const double R = DM_ADD(t2, t1);
    const double hfsq = DM_MUL(DM_MUL(0.5, f), f);
    /* log(x) = k ln2 + f - hfsq + s (hfsq + R), summed smallest-first */
    const double inner = DM_ADD(DM_MUL(s, DM_ADD(hfsq, R)), DM_MUL(dk, LN2_LO));
    return DM_ADD(DM_MUL(dk, LN2_HI), DM_SUB(f, DM_SUB(hfsq, inner)));
}

MCTSB200_HD double dm_kernel_cos(double x, double y) {
    const double C1 = dm_from_bits(0x3fa555555555554cULL);
    const double C2 = dm_from_bits(0xbf56c16c16c15177ULL);
    const double C3 = dm_from_bits(0x3efa01a019cb1590ULL);
    const double C4 = dm_from_bits(0xbe927e4f809c52adULL);
    const double C5 = dm_from_bits(0x3e21ee9ebdb4b1c4ULL);
    const double C6 = dm_from_bits(0xbda8fae9be8838d4ULL);
    const double z = DM_MUL(x, x);
    const double w = DM_MUL(z, z);
    const double ra = DM_MUL(z, DM_ADD(C1, DM_MUL(z, DM_ADD(C2, DM_MUL(z, C3)))));
    const double rb = DM_MUL(DM_MUL(w, w), DM_ADD(C4, DM_MUL(z, DM_ADD(C5, DM_MUL(z, C6)))));
    const double r = DM_ADD(ra, rb);
    const double hz = DM_MUL(0.5, z);
    const double one_m = DM_SUB(1.0, hz);
    return DM_ADD(one_m, DM_ADD(DM_SUB(DM_SUB(1.0, one_m), hz), DM_SUB(DM_MUL(z, r), DM_MUL(x, y))));
}

MCTSB200_HD double dm_kernel_sin(double x, double y) {
    const double S1 = dm_from_bits(0xbfc5555555555549ULL);
    const double S2 = dm_from_bits(0x3f8111111110f8a6ULL);
    const double S3 = dm_from_bits(0xbf2a01a019c161d5ULL);
    const double S4 = dm_from_bits(0x3ec71de357b1fe7dULL);
    const double S5 = dm_from_bits(0xbe5ae5e68a2b9cebULL);
    const double S6 = dm_from_bits(0x3de5d93a5acfd57cULL);
    const double z = DM_MUL(x, x);
    const double w = DM_MUL(z, z);
    const double r = DM_ADD(DM_ADD(S2, DM_MUL(z, DM_ADD(S3, DM_MUL(z, S4)))),
                            DM_MUL(DM_MUL(z, w), DM_ADD(S5, DM_MUL(z, S6))));
    const double v = DM_MUL(z, x);
    /* x - ((z (y/2 - v r) - y) - v S1) */
    const double a = DM_MUL(z, DM_SUB(DM_MUL(0.5, y), DM_MUL(v, r)));
    return DM_SUB(x, DM_SUB(DM_SUB(a, y), DM_MUL(v, S1)));
}

/* Sine / cosine for |x| <= 2^20 (two-stage Cody-Waite reduction by pi/2). Outside that range: NaN. `want_sin` selects the
 * function after the common reduction: sin(y + n pi/2) and cos(y + n pi/2) are +-sin(y) or +-cos(y) by the quadrant n mod 4. */
MCTSB200_HD double det_sincos_impl(double x, int want_sin);
MCTSB200_HD double det_sin(double x) { return det_sincos_impl(x, 1); }
MCTSB200_HD double det_cos(double x) {
    const double INV_PIO2 = dm_from_bits(0x3fe45f306dc9c883ULL);
    const double PIO2_1 = dm_from_bits(0x3ff921fb54400000ULL);  /* first 33 bits of pi/2 */
    const double PIO2_2 = dm_from_bits(0x3dd0b4611a600000ULL);  /* next 33 bits          */
    const double PIO2_2T = dm_from_bits(0x3ba3198a2e037073ULL); /* pi/2 - PIO2_1 - PIO2_2 */
    const uint64_t ax = dm_bits(x) & 0x7fffffffffffffffULL;
    if (ax > 0x4130000000000000ULL) return dm_from_bits(0x7ff8000000000000ULL); /* > 2^20, inf, NaN */
    double fn;
#if defined(__CUDA_ARCH__)
    fn = floor(DM_ADD(DM_MUL(x, INV_PIO2), 0.5));
#else
    fn = floor(DM_ADD(DM_MUL(x, INV_PIO2), 0.5));
#endif
    const int n = (int)fn;
    const double t = DM_SUB(x, DM_MUL(fn, PIO2_1));   /* fn*PIO2_1 exact (33 + 21 bits) */
    const double w1 = DM_MUL(fn, PIO2_2);             /* exact */
    const double r = DM_SUB(t, w1);
    const double w = DM_SUB(DM_MUL(fn, PIO2_2T), DM_SUB(DM_SUB(t, r), w1));
    const double y0 = DM_SUB(r, w);
    const double y1 = DM_SUB(DM_SUB(r, y0), w);
    switch (n & 3) {
        case 0: return dm_kernel_cos(y0, y1);
        case 1: return DM_SUB(0.0, dm_kernel_sin(y0, y1));
        case 2: return DM_SUB(0.0, dm_kernel_cos(y0, y1));
        default: return dm_kernel_sin(y0, y1);
    }
}

MCTSB200_HD double det_sincos_impl(double x, int want_sin) {
    const double INV_PIO2 = dm_from_bits(0x3fe45f306dc9c883ULL);
    const double PIO2_1 = dm_from_bits(0x3ff921fb54400000ULL);
    const double PIO2_2 = dm_from_bits(0x3dd0b4611a600000ULL);
    const double PIO2_2T = dm_from_bits(0x3ba3198a2e037073ULL);
    const uint64_t ax = dm_bits(x) & 0x7fffffffffffffffULL;
    if (ax > 0x4130000000000000ULL) return dm_from_bits(0x7ff8000000000000ULL);
    const double fn = floor(DM_ADD(DM_MUL(x, INV_PIO2), 0.5));
    const int n = (int)fn;
    const double t = DM_SUB(x, DM_MUL(fn, PIO2_1));
    const double w1 = DM_MUL(fn, PIO2_2);
    const double r = DM_SUB(t, w1);
    const double w = DM_SUB(DM_MUL(fn, PIO2_2T), DM_SUB(DM_SUB(t, r), w1));
    const double y0 = DM_SUB(r, w);
    const double y1 = DM_SUB(DM_SUB(r, y0), w);
    switch ((n + (want_sin ? 3 : 0)) & 3) { /* sin(x) = cos(x - pi/2): one quadrant back */
        case 0: return dm_kernel_cos(y0, y1);
        case 1: return DM_SUB(0.0, dm_kernel_sin(y0, y1));
        case 2: return DM_SUB(0.0, dm_kernel_cos(y0, y1));
        default: return dm_kernel_sin(y0, y1);
    }
}

/* Both at once: det_sin(x) and det_cos(x) bit for bit (one reduction, both kernels evaluated on every call and the quadrant applied
 * by selection) -- for callers that need the pair, and for SIMT code whose lanes would otherwise split over the quadrant switch. */
MCTSB200_HD void det_sincos(double x, double* s_out, double* c_out) {
    const double INV_PIO2 = dm_from_bits(0x3fe45f306dc9c883ULL);
    const double PIO2_1 = dm_from_bits(0x3ff921fb54400000ULL);
    const double PIO2_2 = dm_from_bits(0x3dd0b4611a600000ULL);
    const double PIO2_2T = dm_from_bits(0x3ba3198a2e037073ULL);
    const uint64_t ax = dm_bits(x) & 0x7fffffffffffffffULL;
    if (ax > 0x4130000000000000ULL) {
        *s_out = *c_out = dm_from_bits(0x7ff8000000000000ULL);
        return;
    }
    const double fn = floor(DM_ADD(DM_MUL(x, INV_PIO2), 0.5));
    const int n = (int)fn;
    const double t = DM_SUB(x, DM_MUL(fn, PIO2_1));
    const double w1 = DM_MUL(fn, PIO2_2);
    const double r = DM_SUB(t, w1);
    const double w = DM_SUB(DM_MUL(fn, PIO2_2T), DM_SUB(DM_SUB(t, r), w1));
    const double y0 = DM_SUB(r, w);
    const double y1 = DM_SUB(DM_SUB(r, y0), w);
    const double kc = dm_kernel_cos(y0, y1), ks = dm_kernel_sin(y0, y1);
    const double nkc = DM_SUB(0.0, kc), nks = DM_SUB(0.0, ks);
    const int q = n & 3;
    *c_out = q == 0 ? kc : q == 1 ? nks : q == 2 ? nkc : ks;  /* det_cos's switch                    */
    *s_out = q == 0 ? ks : q == 1 ? kc : q == 2 ? nks : nkc;  /* det_sin's: the same, one quadrant back */
}

#endif /* MCTSB200_DETMATH_H */
