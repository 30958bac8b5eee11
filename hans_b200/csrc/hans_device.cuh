// hans_device.cuh -- device-side arithmetic of the walker engine (sm_100a).
//
// Everything here is float64 and written so that the sequence of IEEE operations is fixed by the
// source: the translation unit is compiled with -fmad=false and every fused multiply-add is an
// explicit hd_fma().  That is what makes accept/reject decisions and post-move state bit-identical
// to the CPU checker in tests/ on the same proposals.
//
// Reference call sites replaced (all under /root/reference): the alk.* entry points used by
// MC_run (NesSa/MCNS.py:276-392) and the Python-specified cell moves / predicates
// (MCNS.py:97-252).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/hans_b200.h"
#include "../../include/hans_detmath.h"

namespace hb {

constexpr unsigned FULL = 0xffffffffu;
constexpr uint32_t STREAM_MOVE = 0u;
constexpr uint32_t STREAM_GROW = 1u;

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based generator (replaces alk.random_uniform_random MCNS.py:148,222,308
// and numpy's global stream MCNS.py:174,227,313,351,372).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t (&out)[4])
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        if (r) { k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// two uniforms of block `blk` of the stream keyed (seed; c0, c1, c2)
__device__ __forceinline__ void draw_block(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2,
                                           uint32_t blk, uint32_t stream, double &ua, double &ub)
{
    uint32_t x[4];
    philox4x32_10(c0, c1, c2, blk | (stream << 16), (uint32_t)seed, (uint32_t)(seed >> 32), x);
    ua = hd_u01(x[0], x[1]);
    ub = hd_u01(x[2], x[3]);
}

// The four blocks (blk = 0..3, STREAM_MOVE) of one move at once with the ten rounds ROLLED: the
// same integers as four draw_block() calls in 0.5 KB of code instead of 4.5 KB (the walk kernels are
// instruction-cache bound), and the four blocks give the loop body its instruction-level parallelism.
__device__ __forceinline__ void draw_move_blocks(uint64_t seed, uint32_t lo, uint32_t hi, uint32_t gid, double (&u)[8])
{
    uint32_t c0[4], c1[4], c2[4], c3[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) { c0[b] = lo; c1[b] = hi; c2[b] = gid; c3[b] = (uint32_t)b | (STREAM_MOVE << 16); }
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll 1
    for (int r = 0; r < 10; ++r) {
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, c0[b]), lo0 = 0xD2511F53u * c0[b];
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2[b]), lo1 = 0xCD9E8D57u * c2[b];
            const uint32_t n0 = hi1 ^ c1[b] ^ k0, n2 = hi0 ^ c3[b] ^ k1;
            c0[b] = n0; c1[b] = lo1; c2[b] = n2; c3[b] = lo0;
        }
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
#pragma unroll
    for (int b = 0; b < 4; ++b) { u[2 * b] = hd_u01(c0[b], c1[b]); u[2 * b + 1] = hd_u01(c2[b], c3[b]); }
}

// ---------------------------------------------------------------------------------------------
// Cell algebra: H rows are the cell vectors (MCNS.py:112-113); r = s.H, s = r.H^-1.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double det3(const double *H)
{
    const double c00 = H[4] * H[8] - H[5] * H[7];
    const double c01 = H[5] * H[6] - H[3] * H[8];
    const double c02 = H[3] * H[7] - H[4] * H[6];
    return (H[0] * c00 + H[1] * c01) + H[2] * c02;
}

__device__ __forceinline__ void hinv(const double *H, double *Hi) // inline: the matrices stay in registers
{
    const double c00 = H[4] * H[8] - H[5] * H[7];
    const double c01 = H[5] * H[6] - H[3] * H[8];
    const double c02 = H[3] * H[7] - H[4] * H[6];
    const double det = (H[0] * c00 + H[1] * c01) + H[2] * c02;
    const double rd = 1.0 / det;
    Hi[0] = c00 * rd;
    Hi[3] = c01 * rd;
    Hi[6] = c02 * rd;
    Hi[1] = (H[2] * H[7] - H[1] * H[8]) * rd;
    Hi[4] = (H[0] * H[8] - H[2] * H[6]) * rd;
    Hi[7] = (H[1] * H[6] - H[0] * H[7]) * rd;
    Hi[2] = (H[1] * H[5] - H[2] * H[4]) * rd;
    Hi[5] = (H[2] * H[3] - H[0] * H[5]) * rd;
    Hi[8] = (H[0] * H[4] - H[1] * H[3]) * rd;
}

// o = v.M (row vector times matrix)
__device__ __forceinline__ void vecmat(double v0, double v1, double v2, const double *M, double &o0,
                                       double &o1, double &o2)
{
    o0 = hd_fma(v2, M[6], hd_fma(v1, M[3], v0 * M[0]));
    o1 = hd_fma(v2, M[7], hd_fma(v1, M[4], v0 * M[1]));
    o2 = hd_fma(v2, M[8], hd_fma(v1, M[5], v0 * M[2]));
}

__device__ __forceinline__ double dot3(const double *a, const double *b)
{
    return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2];
}

// np.dot of two 3-vectors as numpy (OpenBLAS ddot, FMA hardware) evaluates it: the shear path of the
// reference (MCNS.py:161-163) is Python, and tests/test_reference_live_cpu.py compares delta_H with the reference executed live
__device__ __forceinline__ double dot3_blas(const double *a, const double *b)
{
    return hd_fma(a[2], b[2], hd_fma(a[1], b[1], a[0] * b[0]));
}

__device__ __forceinline__ void cross3(const double *a, const double *b, double *o)
{
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

// alk.box_min_aspect_ratio / min_aspect_ratio (MCNS.py:97-118)
__device__ __noinline__ double min_aspect_ratio(const double *H)
{
    const double vol = fabs(det3(H));
    double best = 1.7976931348623157e308;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        double n[3];
        cross3(H + 3 * ((i + 1) % 3), H + 3 * ((i + 2) % 3), n);
        const double nn = sqrt(dot3(n, n));
        n[0] /= nn; n[1] /= nn; n[2] /= nn;
        const double w = fabs(dot3(n, H + 3 * i));
        if (w < best) best = w;
    }
    return best / hd_cbrt(vol);
}

// largest |cos| between two cell vectors; min_angle (MCNS.py:120-133) = arccos of it
__device__ __noinline__ double max_abs_cos(const double *H)
{
    double best = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double *a = H + 3 * ((i + 1) % 3);
        const double *b = H + 3 * ((i + 2) % 3);
        double c = dot3(a, b) / sqrt(dot3(a, a) * dot3(b, b));
        if (c < 0) c *= -1;
        if (c > best) best = c;
    }
    return best;
}

// The cell-shape rejections of box_shear_step / box_stretch_step (MCNS.py:194-197,243-246) as
// predicates without cbrt / division / arccos:
//   min_i w_i / V^(1/3) < L,  w_i = V / |a_j x a_k|   <=>   V^2 < (L max_i |a_j x a_k|)^3
//   min angle < A   <=>   max (a.b)^2 / ((a.a)(b.b)) > cos^2 A
__device__ __forceinline__ bool aspect_fails(const double *H, double min_ar)
{
    const double vol = fabs(det3(H));
    double m2 = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        double n[3];
        cross3(H + 3 * ((i + 1) % 3), H + 3 * ((i + 2) % 3), n);
        const double nn = dot3(n, n);
        if (nn > m2) m2 = nn;
    }
    const double t = min_ar * sqrt(m2);
    return (vol * vol) < (t * t) * t;
}

__device__ __forceinline__ bool angle_fails(const double *H, double cos_min_ang)
{
    const double c2 = cos_min_ang * cos_min_ang;
    bool bad = false;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double *a = H + 3 * ((i + 1) % 3);
        const double *b = H + 3 * ((i + 2) % 3);
        const double ab = dot3(a, b);
        bad = bad || (ab * ab > c2 * (dot3(a, a) * dot3(b, b)));
    }
    return bad;
}

// squared minimum-image separation of a Cartesian difference (dx,dy,dz); 30 FP64 operations
__device__ __forceinline__ double pair_r2(const double *H, const double *Hi, double dx, double dy,
                                          double dz)
{
    double s0, s1, s2, m0, m1, m2;
    vecmat(dx, dy, dz, Hi, s0, s1, s2);
    s0 -= hd_rint(s0);
    s1 -= hd_rint(s1);
    s2 -= hd_rint(s2);
    vecmat(s0, s1, s2, H, m0, m1, m2);
    return hd_fma(m2, m2, hd_fma(m1, m1, m0 * m0));
}

// v' = q v q* for a unit quaternion (w,x,y,z)
__device__ __forceinline__ void quat_rot(const double *q, const double *v, double *o)
{
    double t[3], u[3];
    cross3(q + 1, v, t);
    t[0] *= 2.0; t[1] *= 2.0; t[2] *= 2.0;
    cross3(q + 1, t, u);
    o[0] = (v[0] + q[0] * t[0]) + u[0];
    o[1] = (v[1] + q[0] * t[1]) + u[1];
    o[2] = (v[2] + q[0] * t[2]) + u[2];
}

__device__ __forceinline__ void box_muller(double ua, double ub, double &z1, double &z2)
{
    const double rad = sqrt(-2.0 * hd_log(1.0 - ua));
    double s, co;
    hd_sincos(HD_TWO_PI * ub, &s, &co);
    z1 = rad * co;
    z2 = rad * s;
}

// ---------------------------------------------------------------------------------------------
// One proposal (registers).  Draw order of the self-driving mode, SURVEY.md 10.1: per move four
// Philox blocks keyed (seed; ctr, walker id): blk0 = {u_chain, u_type}, blk1 = {u1,u2},
// blk2 = {u3,u4}, blk3 = {u_acc, spare}.
// ---------------------------------------------------------------------------------------------
struct Prop {
    int type, ichain, idx0, idx1;
    double p[4];
    double u_acc;
};

__device__ __noinline__ Prop gen_proposal(const hans_cfg *cp, uint64_t seed, uint32_t gid, uint64_t ctr)
{
    const hans_cfg &c = *cp;
    // every lane generates the proposal of a DIFFERENT move (ctr differs per lane): the scalar work
    // of a batch of moves is spread over the lanes of the walker's group
    const uint32_t lo = (uint32_t)ctr, hi = (uint32_t)(ctr >> 32);
    double u[8];
    draw_move_blocks(seed, lo, hi, gid, u);
    const double u_chain = u[0], u_type = u[1], u1 = u[2], u2 = u[3], u3 = u[4], u4 = u[5], u_acc = u[6];
    Prop o;
    o.u_acc = u_acc;
    o.p[0] = o.p[1] = o.p[2] = o.p[3] = 0.0;
    o.idx0 = o.idx1 = 0;

    int ichain = (int)floor(u_chain * c.nchains); // MCNS.py:308
    if (ichain > c.nchains - 1) ichain = c.nchains - 1;
    o.ichain = ichain;

    int type = 5; // MCNS.py:313-344
#pragma unroll
    for (int k = 5; k >= 0; --k)
        if (u_type < c.move_cum[k]) type = k;
    // first k with u_type < cum[k]: cum is non-decreasing, so the smallest satisfying k wins
    o.type = type;

    switch (type) {
    case HANS_VOL:
        o.p[0] = c.dv_max * (2.0 * u1 - 1.0);
        break;
    case HANS_TRANS:
        o.p[0] = c.dr_max * (2.0 * u1 - 1.0);
        o.p[1] = c.dr_max * (2.0 * u2 - 1.0);
        o.p[2] = c.dr_max * (2.0 * u3 - 1.0);
        break;
    case HANS_ROT: {
        double ax = 2.0 * u1 - 1.0, ay = 2.0 * u2 - 1.0, az = 2.0 * u3 - 1.0;
        const double n2 = (ax * ax + ay * ay) + az * az;
        if (n2 == 0.0) { ax = 1.0; ay = 0.0; az = 0.0; }
        else { const double n = sqrt(n2); ax /= n; ay /= n; az /= n; }
        const double amp = c.dt_max < HD_PI ? c.dt_max : HD_PI;
        const double th = amp * (2.0 * u4 - 1.0);
        double s, co;
        hd_sincos(0.5 * th, &s, &co);
        o.p[0] = co; o.p[1] = s * ax; o.p[2] = s * ay; o.p[3] = s * az;
        break;
    }
    case HANS_DIH: {
        if (c.nbeads < 4) { o.idx0 = -1; break; }
        int ia = (int)floor(u1 * (c.nbeads - 3));
        if (ia > c.nbeads - 4) ia = c.nbeads - 4;
        o.idx0 = 3 + ia;
        o.idx1 = (c.allow_flip && u2 < 0.5) ? 1 : 0;
        const double amp = c.dh_max < HD_PI ? c.dh_max : HD_PI;
        o.p[0] = amp * (2.0 * u3 - 1.0);
        break;
    }
    case HANS_SHEAR: {
        int k = (int)floor(u1 * 3.0);
        if (k > 2) k = 2;
        o.idx0 = k;
        double z1, z2;
        box_muller(u2, u3, z1, z2);
        o.p[0] = c.dshear * z1;
        o.p[1] = c.dshear * z2;
        break;
    }
    default: {
        int i = (int)floor(u1 * 3.0), j = (int)floor(u2 * 3.0);
        if (i > 2) i = 2;
        if (j > 2) j = 2;
        if (i == j) j = (j + 1) % 3;
        o.idx0 = i; o.idx1 = j;
        double z1, z2;
        box_muller(u3, u4, z1, z2);
        o.p[0] = c.dstretch * z1;
        break;
    }
    }
    return o;
}

} // namespace hb
