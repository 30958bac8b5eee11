/*
 * hans_detmath.h -- deterministic elementary functions for the hans-b200 walker engine.
 *
 * Why this exists: the self-driving MC walk needs exp/log/sin/cos/cbrt (Gaussian shear and
 * stretch magnitudes, MCNS.py:174-175,227,231-232; the volume-move weight exp(n ln V'/V),
 * MCNS.py:266; the unit-volume scaling cbrt(V), MCNS.py:118).  glibc's libm and CUDA's libdevice
 * round these differently in the last bit, which would make "GPU trajectory == CPU oracle
 * trajectory" impossible.  Everything below is built only from IEEE-754 binary64
 * add / multiply / divide / sqrt / fma, which are correctly rounded on both x86-64 and sm_100a,
 * so the same source gives bit-identical results on both sides PROVIDED the translation unit is
 * compiled without implicit contraction (gcc: -ffp-contract=off, nvcc: -fmad=false); fused
 * operations are written explicitly as hd_fma().
 *
 * Accuracy target: a few ulp (checked against libm in tests/test_detmath.py).  These are not
 * correctly-rounded libm replacements and are not meant to be.
 *
 * Included by BOTH the CUDA product (hans_b200/csrc) and the CPU oracle (oracle/); it contains
 * no algorithm of the reference path, only arithmetic primitives.
 */
#ifndef HANS_DETMATH_H
#define HANS_DETMATH_H

#include <stdint.h>
#include <string.h>
#include <math.h>

#if defined(__CUDACC__)
#define HANS_HD __host__ __device__ __forceinline__
/* the transcendental-sized routines are kept out of line on the GPU: one copy in the instruction
 * cache instead of one per call site (they are off the pair-test hot loop) */
#define HANS_HD_BIG __host__ __device__ __noinline__
#else
#define HANS_HD static inline
#define HANS_HD_BIG static inline
#endif

#define HD_PI 3.14159265358979323846264338327950288
#define HD_TWO_PI 6.28318530717958647692528676655900577

HANS_HD double hd_fma(double a, double b, double c)
{
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}

HANS_HD uint64_t hd_bits(double x)
{
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(x);
#else
    uint64_t u;
    memcpy(&u, &x, sizeof u);
    return u;
#endif
}

HANS_HD double hd_from_bits(uint64_t u)
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double x;
    memcpy(&x, &u, sizeof x);
    return x;
#endif
}

/* Round to nearest integer, ties to even, valid for |x| < 2^51 (the engine never exceeds that).
 * Identical to rint() under the default rounding mode but stays on the FP64 add pipe on the GPU. */
HANS_HD double hd_rint(double x)
{
    const double magic = 6755399441055744.0; /* 1.5 * 2^52 */
    double t = x + magic; /* not foldable without -ffast-math / -fassociative-math */
    return t - magic;
}

/* 2^k for -1022 <= k <= 1023 */
HANS_HD double hd_pow2i(int k)
{
    return hd_from_bits((uint64_t)(k + 1023) << 52);
}

/* exp(x).  Cody-Waite reduction x = k ln2 + r, |r| <= ln2/2, Taylor to r^13. */
HANS_HD_BIG double hd_exp(double x)
{
    if (x != x) return x;
    if (x > 709.0) return hd_from_bits(0x7ff0000000000000ull);
    if (x < -708.0) return 0.0;
    const double inv_ln2 = 1.44269504088896338700e+00;
    const double ln2_hi = 6.93147180369123816490e-01;
    const double ln2_lo = 1.90821492927058770002e-10;
    double k = hd_rint(x * inv_ln2);
    double r = hd_fma(-k, ln2_hi, x);
    r = hd_fma(-k, ln2_lo, r);
    double p = 1.0 / 6227020800.0; /* 1/13! */
    p = hd_fma(p, r, 1.0 / 479001600.0);
    p = hd_fma(p, r, 1.0 / 39916800.0);
    p = hd_fma(p, r, 1.0 / 3628800.0);
    p = hd_fma(p, r, 1.0 / 362880.0);
    p = hd_fma(p, r, 1.0 / 40320.0);
    p = hd_fma(p, r, 1.0 / 5040.0);
    p = hd_fma(p, r, 1.0 / 720.0);
    p = hd_fma(p, r, 1.0 / 120.0);
    p = hd_fma(p, r, 1.0 / 24.0);
    p = hd_fma(p, r, 1.0 / 6.0);
    p = hd_fma(p, r, 0.5);
    p = hd_fma(p, r, 1.0);
    p = hd_fma(p, r, 1.0);
    int ki = (int)k;
    /* split the scaling so that 2^ki never leaves the normal range */
    int k1 = ki / 2, k2 = ki - k1;
    return (p * hd_pow2i(k1)) * hd_pow2i(k2);
}

/* log(x) for finite x > 0 (normal numbers; subnormals are scaled up first). */
HANS_HD_BIG double hd_log(double x)
{
    if (x != x) return x;
    if (x <= 0.0) return (x == 0.0) ? -hd_from_bits(0x7ff0000000000000ull)
                                   : hd_from_bits(0x7ff8000000000000ull);
    int e = 0;
    if (x < 2.2250738585072014e-308) { x *= 18014398509481984.0; e = -54; } /* 2^54 */
    uint64_t b = hd_bits(x);
    e += (int)((b >> 52) & 0x7ff) - 1023;
    double m = hd_from_bits((b & 0x000fffffffffffffull) | 0x3ff0000000000000ull); /* [1,2) */
    if (m > 1.41421356237309504880) { m *= 0.5; e += 1; }
    double f = m - 1.0;
    double s = f / (2.0 + f);
    double z = s * s;
    double p = 1.0 / 25.0;
    p = hd_fma(p, z, 1.0 / 23.0);
    p = hd_fma(p, z, 1.0 / 21.0);
    p = hd_fma(p, z, 1.0 / 19.0);
    p = hd_fma(p, z, 1.0 / 17.0);
    p = hd_fma(p, z, 1.0 / 15.0);
    p = hd_fma(p, z, 1.0 / 13.0);
    p = hd_fma(p, z, 1.0 / 11.0);
    p = hd_fma(p, z, 1.0 / 9.0);
    p = hd_fma(p, z, 1.0 / 7.0);
    p = hd_fma(p, z, 1.0 / 5.0);
    p = hd_fma(p, z, 1.0 / 3.0);
    /* log m = 2s + 2s*z*p */
    double two_s = 2.0 * s;
    double logm = hd_fma(two_s * z, p, two_s);
    const double ln2_hi = 6.93147180369123816490e-01;
    const double ln2_lo = 1.90821492927058770002e-10;
    double de = (double)e;
    return hd_fma(de, ln2_hi, hd_fma(de, ln2_lo, logm));
}

/* sin and cos of x, intended for |x| <= ~1e5 (the engine clamps rotation amplitudes to pi). */
HANS_HD_BIG void hd_sincos(double x, double *sn, double *cs)
{
    const double two_over_pi = 6.36619772367581382433e-01;
    const double pio2_1 = 1.57079632673412561417e+00;  /* first 33 bits of pi/2 */
    const double pio2_2 = 6.07710050630396597660e-11;  /* next 33 bits */
    const double pio2_3 = 2.02226624879595063154e-21;  /* the rest */
    double k = hd_rint(x * two_over_pi);
    double r = hd_fma(-k, pio2_1, x);
    r = hd_fma(-k, pio2_2, r);
    r = hd_fma(-k, pio2_3, r);
    double z = r * r;
    /* sin r = r (1 - z/3! + z^2/5! - ... - z^9/19!) */
    double ps = -1.0 / 121645100408832000.0;          /* -1/19! */
    ps = hd_fma(ps, z, 1.0 / 355687428096000.0);      /* 1/17! */
    ps = hd_fma(ps, z, -1.0 / 1307674368000.0);       /* -1/15! */
    ps = hd_fma(ps, z, 1.0 / 6227020800.0);           /* 1/13! */
    ps = hd_fma(ps, z, -1.0 / 39916800.0);            /* -1/11! */
    ps = hd_fma(ps, z, 1.0 / 362880.0);               /* 1/9! */
    ps = hd_fma(ps, z, -1.0 / 5040.0);                /* -1/7! */
    ps = hd_fma(ps, z, 1.0 / 120.0);                  /* 1/5! */
    ps = hd_fma(ps, z, -1.0 / 6.0);                   /* -1/3! */
    double s = hd_fma(r * z, ps, r);
    /* cos r = 1 - z/2 + z^2/4! - ... + z^9/18! */
    double pc = -1.0 / 6402373705728000.0;            /* -1/18! */
    pc = hd_fma(pc, z, 1.0 / 20922789888000.0);       /* 1/16! */
    pc = hd_fma(pc, z, -1.0 / 87178291200.0);         /* -1/14! */
    pc = hd_fma(pc, z, 1.0 / 479001600.0);            /* 1/12! */
    pc = hd_fma(pc, z, -1.0 / 3628800.0);             /* -1/10! */
    pc = hd_fma(pc, z, 1.0 / 40320.0);                /* 1/8! */
    pc = hd_fma(pc, z, -1.0 / 720.0);                 /* -1/6! */
    pc = hd_fma(pc, z, 1.0 / 24.0);                   /* 1/4! */
    pc = hd_fma(pc, z, -0.5);                         /* -1/2! */
    double c = hd_fma(z, pc, 1.0);
    long long q = (long long)k;
    switch ((int)(q & 3)) {
    case 0: *sn = s;  *cs = c;  break;
    case 1: *sn = c;  *cs = -s; break;
    case 2: *sn = -s; *cs = -c; break;
    default: *sn = -c; *cs = s; break;
    }
}

/* cbrt(x) for finite x > 0: bit-trick initial guess (a few percent), then Halley iterations
 * y <- y (y^3 + 2x) / (2 y^3 + x) (cubic convergence) and one Newton polish. */
HANS_HD_BIG double hd_cbrt(double x)
{
    if (!(x > 0.0)) return (x == 0.0) ? 0.0 : -hd_from_bits(0x7ff8000000000000ull);
    double scale = 1.0;
    if (x < 2.2250738585072014e-308) { x *= 18014398509481984.0; scale = 3.814697265625e-06; } /* 2^54, 2^-18 */
    double y = hd_from_bits(hd_bits(x) / 3ull + 0x2A9F7893782DA1CEull);
    for (int it = 0; it < 3; ++it) {
        double y3 = (y * y) * y;
        y = y * ((y3 + 2.0 * x) / (2.0 * y3 + x));
    }
    y = y - ((y * y) * y - x) / (3.0 * (y * y));
    return y * scale;
}

/* x^n for an integer n >= 0 by binary exponentiation (fixed multiplication order) */
HANS_HD double hd_powi(double x, int n)
{
    double r = 1.0, b = x;
    while (n > 0) {
        if (n & 1) r = r * b;
        n >>= 1;
        if (n) b = b * b;
    }
    return r;
}

/* 53-bit uniform in [0,1) from two 32-bit words (hi supplies the top bits). */
HANS_HD double hd_u01(uint32_t hi, uint32_t lo)
{
    uint64_t v = ((uint64_t)hi << 32) | (uint64_t)lo;
    return (double)(v >> 11) * 1.1102230246251565404e-16; /* 2^-53 */
}

#endif /* HANS_DETMATH_H */
