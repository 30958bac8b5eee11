/*
 * hans_oracle.c -- CPU oracle (float64, scalar) for the hans nested-sampling MC walker path.
 * TEST INFRASTRUCTURE ONLY.  Parity: pinned to the reference's own Python executed live (tests/test_reference_live_cpu.py),
 * unpinned for hs_alkane's Fortran internals [A1..A10] -- see hans_oracle.h.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off; fused operations are explicit hd_fma()).
 */
#include "hans_oracle.h"
#include "../include/hans_detmath.h"

#include <float.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define HO_MAX_BEADS 64 /* beads per chain the stack buffers are sized for */

/* ---------------------------------------------------------------------------------------------
 * Philox4x32-10 (Salmon et al., SC'11; Random123).  Replaces hs_alkane's generator
 * (alk.random_uniform_random, MCNS.py:148,222,308) and numpy's global stream
 * (MCNS.py:174,227,313,351,372) [A10]: only distributional equivalence is required.
 * ------------------------------------------------------------------------------------------- */
void ho_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        if (r) { k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

double ho_det_exp(double x) { return hd_exp(x); }
double ho_det_log(double x) { return hd_log(x); }
double ho_det_cbrt(double x) { return hd_cbrt(x); }
void ho_det_sincos(double x, double *s, double *c) { hd_sincos(x, s, c); }
double ho_det_rint(double x) { return hd_rint(x); }

/* stream ids packed in counter word 3 */
#define HO_STREAM_MOVE 0u
#define HO_STREAM_GROW 1u

static void draw_block(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t blk,
                       uint32_t stream, double u[2])
{
    uint32_t ctr[4] = { c0, c1, c2, blk | (stream << 16) };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    uint32_t x[4];
    ho_philox4x32_10(ctr, key, x);
    u[0] = hd_u01(x[0], x[1]);
    u[1] = hd_u01(x[2], x[3]);
}

/* ---------------------------------------------------------------------------------------------
 * Cell algebra.  H rows = cell vectors.
 * ------------------------------------------------------------------------------------------- */
static double det3(const double *H)
{
    double c00 = H[4] * H[8] - H[5] * H[7];
    double c01 = H[5] * H[6] - H[3] * H[8];
    double c02 = H[3] * H[7] - H[4] * H[6];
    return (H[0] * c00 + H[1] * c01) + H[2] * c02;
}

/* inverse by cofactors; the box's reciprocal matrix in hs_alkane (2*pi scaling dropped, [A1]) */
void ho_hinv(const double *H, double *Hi)
{
    double c00 = H[4] * H[8] - H[5] * H[7];
    double c01 = H[5] * H[6] - H[3] * H[8];
    double c02 = H[3] * H[7] - H[4] * H[6];
    double det = (H[0] * c00 + H[1] * c01) + H[2] * c02;
    double rd = 1.0 / det;
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

/* alk.box_compute_volume (MCNS.py:105,350,392; mpihans.py:183) */
double ho_volume(const double *H) { return fabs(det3(H)); }

/* v.M for a row vector v */
static void vecmat(const double *v, const double *M, double *o)
{
    o[0] = hd_fma(v[2], M[6], hd_fma(v[1], M[3], v[0] * M[0]));
    o[1] = hd_fma(v[2], M[7], hd_fma(v[1], M[4], v[0] * M[1]));
    o[2] = hd_fma(v[2], M[8], hd_fma(v[1], M[5], v[0] * M[2]));
}

static double dot3(const double *a, const double *b)
{
    return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2];
}

/* np.dot of two 3-vectors as numpy evaluates it in this image (OpenBLAS ddot on an FMA machine):
 * fma(a2,b2, fma(a1,b1, a0*b0)) -- established by tests/test_reference_live_cpu.py against the
 * reference executed here.  Used where the reference's Python calls np.dot on the shear path
 * (MCNS.py:161-163), so that delta_H is bit-identical to the reference's. */
static double dot3_blas(const double *a, const double *b)
{
    return hd_fma(a[2], b[2], hd_fma(a[1], b[1], a[0] * b[0]));
}

static void cross3(const double *a, const double *b, double *o)
{
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

/* min_aspect_ratio, MCNS.py:97-118 (alk.box_min_aspect_ratio at MCNS.py:194,243 is the fork's
 * native version of the same quantity) */
double ho_min_aspect_ratio(const double *H)
{
    double vol = ho_volume(H);
    double best = DBL_MAX;
    for (int i = 0; i < 3; ++i) {
        const double *vi = H + 3 * i;
        double n[3];
        cross3(H + 3 * ((i + 1) % 3), H + 3 * ((i + 2) % 3), n);
        double nn = sqrt(dot3(n, n));
        n[0] /= nn; n[1] /= nn; n[2] /= nn;
        double w = fabs(dot3(n, vi));
        if (w < best) best = w;
    }
    return best / hd_cbrt(vol);
}

/* min_angle, MCNS.py:120-133, stated through the cosine: the smallest angle is the arccos of the
 * largest |cos|; the caller compares against cos(limit) instead of taking arccos. */
double ho_max_abs_cos(const double *H)
{
    double best = 0.0;
    for (int i = 0; i < 3; ++i) {
        const double *a = H + 3 * ((i + 1) % 3);
        const double *b = H + 3 * ((i + 2) % 3);
        double c = dot3(a, b) / sqrt(dot3(a, a) * dot3(b, b));
        if (c < 0) c *= -1;
        if (c > best) best = c;
    }
    return best;
}

/* The two cell-shape rejections of box_shear_step / box_stretch_step (MCNS.py:194-197,243-246)
 * as PREDICATES, free of cbrt / division / arccos:
 *   min_i w_i / V^(1/3) < L  with  w_i = V / |a_j x a_k|   <=>   V^2 < (L max_i |a_j x a_k|)^3
 *   min angle < A  <=>  max (a.b)^2 / ((a.a)(b.b)) > cos^2 A
 * (alk.box_min_aspect_ratio is the hs_alkane fork's native routine: its last-bit arithmetic is
 * unpinned anyway; ho_min_aspect_ratio / ho_max_abs_cos above return the values.) */
int ho_aspect_fails(const double *H, double min_ar)
{
    double vol = ho_volume(H);
    double m2 = 0.0;
    for (int i = 0; i < 3; ++i) {
        double n[3];
        cross3(H + 3 * ((i + 1) % 3), H + 3 * ((i + 2) % 3), n);
        double nn = dot3(n, n);
        if (nn > m2) m2 = nn;
    }
    double t = min_ar * sqrt(m2);
    return (vol * vol) < (t * t) * t;
}

int ho_angle_fails(const double *H, double cos_min_ang)
{
    double c2 = cos_min_ang * cos_min_ang;
    for (int i = 0; i < 3; ++i) {
        const double *a = H + 3 * ((i + 1) % 3);
        const double *b = H + 3 * ((i + 2) % 3);
        double ab = dot3(a, b);
        if (ab * ab > c2 * (dot3(a, a) * dot3(b, b))) return 1;
    }
    return 0;
}

/* squared minimum-image separation [A1]: fractional difference through H^-1, subtract the
 * nearest integer, map back through H.  Exact for the overlap decision whenever every
 * perpendicular cell width exceeds 2 sigma. */
double ho_pair_r2(const double *H, const double *Hi, const double *ri, const double *rj)
{
    double d[3] = { rj[0] - ri[0], rj[1] - ri[1], rj[2] - ri[2] };
    double s[3], m[3];
    vecmat(d, Hi, s);
    s[0] -= rint(s[0]);
    s[1] -= rint(s[1]);
    s[2] -= rint(s[2]);
    vecmat(s, H, m);
    return hd_fma(m[2], m[2], hd_fma(m[1], m[1], m[0] * m[0]));
}

/* brute-force 27-image minimum (test helper: shows when the single-image rule is exact) */
double ho_pair_r2_27(const double *H, const double *ri, const double *rj)
{
    double Hi[9];
    ho_hinv(H, Hi);
    double d[3] = { rj[0] - ri[0], rj[1] - ri[1], rj[2] - ri[2] };
    double s[3];
    vecmat(d, Hi, s);
    s[0] -= rint(s[0]); s[1] -= rint(s[1]); s[2] -= rint(s[2]);
    double best = DBL_MAX;
    for (int a = -1; a <= 1; ++a)
        for (int b = -1; b <= 1; ++b)
            for (int c = -1; c <= 1; ++c) {
                double t[3] = { s[0] + a, s[1] + b, s[2] + c }, m[3];
                vecmat(t, H, m);
                double r2 = dot3(m, m);
                if (r2 < best) best = r2;
            }
    return best;
}

/* MCNS.py:296-301 */
int ho_moves_per_sweep(int nbeads, int nchains)
{
    if (nbeads == 1) return nchains + 7;
    if (nbeads <= 3) return 2 * nchains + 7;
    return (nbeads - 1) * nchains + 7;
}

/* ---------------------------------------------------------------------------------------------
 * Overlap predicates [A2]: overlap iff r^2 < sigma^2, sigma = 1.
 * ------------------------------------------------------------------------------------------- */
#define R_AT(R, c, ic, ib) ((R) + 3 * ((size_t)(ic) * (c)->nbeads + (ib)))

/* beads [b0,nbeads) of `chain` (coordinates in `x`) against every bead of every other chain */
static int inter_overlap(const ho_cfg *c, const double *H, const double *Hi, const double *R,
                         int ichain, const double *x, int b0)
{
    for (int jc = 0; jc < c->nchains; ++jc) {
        if (jc == ichain) continue;
        for (int jb = 0; jb < c->nbeads; ++jb) {
            const double *rj = R_AT(R, c, jc, jb);
            for (int ib = b0; ib < c->nbeads; ++ib)
                if (ho_pair_r2(H, Hi, x + 3 * ib, rj) < 1.0) return 1;
        }
    }
    return 0;
}

/* pairs (j, b) of one chain with j < b, b >= b0, b - j > nexclude [A3] */
static int intra_overlap(const ho_cfg *c, const double *H, const double *Hi, const double *x,
                         int b0)
{
    for (int b = b0; b < c->nbeads; ++b)
        for (int j = 0; j < b - c->nexclude; ++j)
            if (ho_pair_r2(H, Hi, x + 3 * j, x + 3 * b) < 1.0) return 1;
    return 0;
}

/* alk.alkane_chain_inter_boltz equivalent as a predicate */
int ho_chain_overlap(const ho_cfg *c, const double *H, const double *R, int ichain)
{
    double Hi[9];
    ho_hinv(H, Hi);
    return inter_overlap(c, H, Hi, R, ichain, R_AT(R, c, ichain, 0), 0);
}

static int overlap_all(const ho_cfg *c, const double *H, const double *Hi, const double *R,
                       int with_intra)
{
    for (int ic = 0; ic < c->nchains; ++ic) {
        const double *x = R_AT(R, c, ic, 0);
        /* inter: only chains jc > ic, every pair once */
        for (int jc = ic + 1; jc < c->nchains; ++jc)
            for (int jb = 0; jb < c->nbeads; ++jb) {
                const double *rj = R_AT(R, c, jc, jb);
                for (int ib = 0; ib < c->nbeads; ++ib)
                    if (ho_pair_r2(H, Hi, x + 3 * ib, rj) < 1.0) return 1;
            }
        if (with_intra && intra_overlap(c, H, Hi, x, 0)) return 1;
    }
    return 0;
}

/* alk.alkane_check_chain_overlap (MCNS.py:198,247,543) */
int ho_check_overlap(const ho_cfg *c, const double *H, const double *R)
{
    double Hi[9];
    ho_hinv(H, Hi);
    return overlap_all(c, H, Hi, R, 1);
}

int ho_check_overlap_inter(const ho_cfg *c, const double *H, const double *R)
{
    double Hi[9];
    ho_hinv(H, Hi);
    return overlap_all(c, H, Hi, R, 0);
}

/* ---------------------------------------------------------------------------------------------
 * Box change: chains follow the cell rigidly, anchored on their first bead [A6][A7].
 * ------------------------------------------------------------------------------------------- */
static void follow_cell(const ho_cfg *c, const double *Hi_old, const double *H_new, double *R)
{
    for (int ic = 0; ic < c->nchains; ++ic) {
        double *a = R_AT(R, c, ic, 0);
        double f[3], an[3], sh[3];
        vecmat(a, Hi_old, f);
        vecmat(f, H_new, an);
        sh[0] = an[0] - a[0]; sh[1] = an[1] - a[1]; sh[2] = an[2] - a[2];
        for (int ib = 0; ib < c->nbeads; ++ib) {
            double *r = R_AT(R, c, ic, ib);
            r[0] += sh[0]; r[1] += sh[1]; r[2] += sh[2];
        }
    }
}

/* alk.alkane_change_box(ibox, dH) (MCNS.py:186,238,368) */
void ho_change_box(const ho_cfg *c, double *H, double *R, const double *dH)
{
    double Hi[9], Hn[9];
    ho_hinv(H, Hi);
    for (int k = 0; k < 9; ++k) Hn[k] = H[k] + dH[k];
    follow_cell(c, Hi, Hn, R);
    for (int k = 0; k < 9; ++k) H[k] = Hn[k];
}

/* v' = q v q* for a unit quaternion q = (w, x, y, z) */
static void quat_rot(const double *q, const double *v, double *o)
{
    double t[3], u[3];
    cross3(q + 1, v, t);
    t[0] *= 2.0; t[1] *= 2.0; t[2] *= 2.0;
    cross3(q + 1, t, u);
    o[0] = (v[0] + q[0] * t[0]) + u[0];
    o[1] = (v[1] + q[0] * t[1]) + u[1];
    o[2] = (v[2] + q[0] * t[2]) + u[2];
}

/* ---------------------------------------------------------------------------------------------
 * Proposal generation: the documented draw order of the self-driving mode.
 * Per move, four Philox blocks keyed by (seed; ctr, walker id): blk0 = {u_chain, u_type},
 * blk1 = {u1,u2}, blk2 = {u3,u4}, blk3 = {u_acc, spare}.
 * ------------------------------------------------------------------------------------------- */
static void box_muller(double ua, double ub, double *z1, double *z2)
{
    double rad = sqrt(-2.0 * hd_log(1.0 - ua));
    double s, co;
    hd_sincos(HD_TWO_PI * ub, &s, &co);
    *z1 = rad * co;
    *z2 = rad * s;
}

void ho_gen_proposal(const ho_cfg *c, uint64_t seed, uint32_t gid, uint64_t ctr, ho_proposal *out)
{
    double b0[2], b1[2], b2[2], b3[2];
    uint32_t lo = (uint32_t)ctr, hi = (uint32_t)(ctr >> 32);
    draw_block(seed, lo, hi, gid, 0, HO_STREAM_MOVE, b0);
    draw_block(seed, lo, hi, gid, 1, HO_STREAM_MOVE, b1);
    draw_block(seed, lo, hi, gid, 2, HO_STREAM_MOVE, b2);
    draw_block(seed, lo, hi, gid, 3, HO_STREAM_MOVE, b3);
    double u1 = b1[0], u2 = b1[1], u3 = b2[0], u4 = b2[1];

    out->p[0] = out->p[1] = out->p[2] = out->p[3] = 0.0;
    out->idx0 = out->idx1 = 0;
    out->reserved = 0.0;
    out->u_acc = b3[0];

    /* MCNS.py:308 */
    int ichain = (int)floor(b0[0] * c->nchains);
    if (ichain > c->nchains - 1) ichain = c->nchains - 1;
    out->ichain = ichain;

    /* MCNS.py:313-344 */
    int type = 5;
    for (int k = 0; k < 6; ++k)
        if (b0[1] < c->move_cum[k]) { type = k; break; }
    out->type = type;

    switch (type) {
    case HO_VOL: /* [A6] dV uniform in +-dv_max */
        out->p[0] = c->dv_max * (2.0 * u1 - 1.0);
        break;
    case HO_TRANS: /* [A4] */
        out->p[0] = c->dr_max * (2.0 * u1 - 1.0);
        out->p[1] = c->dr_max * (2.0 * u2 - 1.0);
        out->p[2] = c->dr_max * (2.0 * u3 - 1.0);
        break;
    case HO_ROT: { /* [A5] axis = normalised random vector in the cube, angle +-dt_max */
        double ax = 2.0 * u1 - 1.0, ay = 2.0 * u2 - 1.0, az = 2.0 * u3 - 1.0;
        double n2 = (ax * ax + ay * ay) + az * az;
        if (n2 == 0.0) { ax = 1.0; ay = 0.0; az = 0.0; }
        else { double n = sqrt(n2); ax /= n; ay /= n; az /= n; }
        double amp = c->dt_max < HD_PI ? c->dt_max : HD_PI;
        double th = amp * (2.0 * u4 - 1.0);
        double s, co;
        hd_sincos(0.5 * th, &s, &co);
        out->p[0] = co; out->p[1] = s * ax; out->p[2] = s * ay; out->p[3] = s * az;
        break;
    }
    case HO_DIH: { /* [A8] */
        if (c->nbeads < 4) { out->idx0 = -1; break; }
        int ia = (int)floor(u1 * (c->nbeads - 3));
        if (ia > c->nbeads - 4) ia = c->nbeads - 4;
        out->idx0 = 3 + ia;
        out->idx1 = (c->allow_flip && u2 < 0.5) ? 1 : 0;
        double amp = c->dh_max < HD_PI ? c->dh_max : HD_PI;
        out->p[0] = amp * (2.0 * u3 - 1.0);
        break;
    }
    case HO_SHEAR: { /* MCNS.py:148,174-175 */
        int k = (int)floor(u1 * 3.0);
        if (k > 2) k = 2;
        out->idx0 = k;
        double z1, z2;
        box_muller(u2, u3, &z1, &z2);
        out->p[0] = c->dshear * z1;
        out->p[1] = c->dshear * z2;
        break;
    }
    default: { /* HO_STRETCH, MCNS.py:222-227 */
        int i = (int)floor(u1 * 3.0), j = (int)floor(u2 * 3.0);
        if (i > 2) i = 2;
        if (j > 2) j = 2;
        if (i == j) j = (j + 1) % 3;
        out->idx0 = i; out->idx1 = j;
        double z1, z2;
        box_muller(u3, u4, &z1, &z2);
        out->p[0] = c->dstretch * z1;
        break;
    }
    }
}

/* ---------------------------------------------------------------------------------------------
 * One trial move.  mode HO_MODE_FUSED = MC_run's accept/revert (MCNS.py:349-380);
 * HO_MODE_PROPOSE = the hs_alkane entry point alone (state left at the trial configuration,
 * boltz returned, the caller reverts).
 * ------------------------------------------------------------------------------------------- */
static void set_dec(ho_decision *d, int acc, int reason, double boltz)
{
    d->accepted = acc; d->reason = reason; d->boltz = boltz;
}

/* single-chain moves: alk.alkane_translate_chain / rotate_chain / bond_rotate
 * (MCNS.py:323,328,333), accept/restore MCNS.py:371-380 */
static void move_chain(const ho_cfg *c, double *H, double *R, const ho_proposal *p, int mode,
                       ho_decision *out)
{
    const int nb = c->nbeads;
    double *x = R_AT(R, c, p->ichain, 0);
    double bak[3 * HO_MAX_BEADS];
    for (int k = 0; k < 3 * nb; ++k) bak[k] = x[k];
    int b0 = 0, intra = 0;

    if (p->type == HO_TRANS) {
        for (int b = 0; b < nb; ++b) {
            x[3 * b + 0] += p->p[0]; x[3 * b + 1] += p->p[1]; x[3 * b + 2] += p->p[2];
        }
    } else if (p->type == HO_ROT) {
        double com[3] = { 0.0, 0.0, 0.0 };
        for (int b = 0; b < nb; ++b) {
            com[0] += x[3 * b + 0]; com[1] += x[3 * b + 1]; com[2] += x[3 * b + 2];
        }
        double dn = (double)nb;
        com[0] /= dn; com[1] /= dn; com[2] /= dn;
        for (int b = 0; b < nb; ++b) {
            double v[3] = { x[3 * b] - com[0], x[3 * b + 1] - com[1], x[3 * b + 2] - com[2] }, o[3];
            quat_rot(p->p, v, o);
            x[3 * b + 0] = o[0] + com[0]; x[3 * b + 1] = o[1] + com[1]; x[3 * b + 2] = o[2] + com[2];
        }
    } else { /* HO_DIH */
        int ia = p->idx0;
        if (nb < 4 || ia < 3 || ia > nb - 1) { set_dec(out, 0, HO_REJ_INVALID, 0.0); return; }
        if (p->idx1) { /* reverse the bead order: the other end of the chain becomes the tail */
            for (int b = 0; b < nb / 2; ++b)
                for (int k = 0; k < 3; ++k) {
                    double t = x[3 * b + k];
                    x[3 * b + k] = x[3 * (nb - 1 - b) + k];
                    x[3 * (nb - 1 - b) + k] = t;
                }
        }
        const double *pv = x + 3 * (ia - 1), *pw = x + 3 * (ia - 2);
        double ax[3] = { pv[0] - pw[0], pv[1] - pw[1], pv[2] - pw[2] };
        double n = sqrt(dot3(ax, ax));
        ax[0] /= n; ax[1] /= n; ax[2] /= n;
        double s, co;
        hd_sincos(0.5 * p->p[0], &s, &co);
        double q[4] = { co, s * ax[0], s * ax[1], s * ax[2] };
        double piv[3] = { pv[0], pv[1], pv[2] };
        for (int b = ia; b < nb; ++b) {
            double v[3] = { x[3 * b] - piv[0], x[3 * b + 1] - piv[1], x[3 * b + 2] - piv[2] }, o[3];
            quat_rot(q, v, o);
            x[3 * b + 0] = o[0] + piv[0]; x[3 * b + 1] = o[1] + piv[1]; x[3 * b + 2] = o[2] + piv[2];
        }
        b0 = ia;
        intra = 1;
    }

    double Hi[9];
    ho_hinv(H, Hi);
    int ov = inter_overlap(c, H, Hi, R, p->ichain, x, b0);
    if (!ov && intra) ov = intra_overlap(c, H, Hi, x, b0);
    double boltz = ov ? 0.0 : 1.0;
    if (mode == HO_MODE_PROPOSE) { set_dec(out, !ov, ov ? HO_REJ_OVERLAP : HO_ACCEPT, boltz); return; }
    if (p->u_acc < boltz) { set_dec(out, 1, HO_ACCEPT, boltz); return; } /* MCNS.py:372 */
    for (int k = 0; k < 3 * nb; ++k) x[k] = bak[k];                       /* MCNS.py:379-380 */
    set_dec(out, 0, HO_REJ_OVERLAP, boltz);
}

/* alk.alkane_box_resize(pressure, ibox, 0) + acceptance MCNS.py:349-357 [A6] */
static void move_volume(const ho_cfg *c, double *H, double *R, const ho_proposal *p,
                        double volume_limit, int mode, ho_decision *out)
{
    const size_t n3 = (size_t)3 * c->nchains * c->nbeads;
    double vol = ho_volume(H);
    double vnew_t = vol + p->p[0];
    if (!(vnew_t > 0.0)) { set_dec(out, 0, HO_REJ_INVALID, 0.0); return; }
    /* a proposal clearly above the ceiling is rejected before any further arithmetic: the
     * recomputed volume |det(sH)| differs from V + dV by a few ulp only (MCNS.py:351) */
    if (mode == HO_MODE_FUSED && (vnew_t - volume_limit) > 1e-9 * fabs(volume_limit)) {
        set_dec(out, 0, HO_REJ_CEILING, 0.0);
        return;
    }
    double ratio = vnew_t / vol;
    double s = hd_cbrt(ratio);
    double Hn[9], Hi[9];
    for (int k = 0; k < 9; ++k) Hn[k] = H[k] * s;
    double new_volume = ho_volume(Hn); /* MCNS.py:350 recomputes it from the cell */
    /* acceptance weight exp(-P dV + nchains ln(V'/V)) (mirrored at MCNS.py:266), evaluated as
     * (V'/V)^nchains * exp(-P dV); the exponential is skipped for P = 0 (always, from mpihans.py) */
    double boltz_f = hd_powi(ratio, c->nchains);
    if (c->pressure != 0.0) boltz_f = boltz_f * hd_exp(-c->pressure * p->p[0]);

    if (mode == HO_MODE_FUSED) {
        /* The reference evaluates overlap -> u_acc -> ceiling; the conjunction is order
         * independent, so the O(N^2) test is evaluated last here. */
        if (!((new_volume - volume_limit) < DBL_EPSILON)) { set_dec(out, 0, HO_REJ_CEILING, boltz_f); return; }
        if (!(p->u_acc < boltz_f)) { set_dec(out, 0, HO_REJ_BOLTZ, boltz_f); return; }
    }
    double *bak = (double *)malloc(sizeof(double) * (n3 + 9));
    for (size_t k = 0; k < n3; ++k) bak[k] = R[k];
    for (int k = 0; k < 9; ++k) bak[n3 + k] = H[k];
    ho_hinv(H, Hi);
    follow_cell(c, Hi, Hn, R);
    for (int k = 0; k < 9; ++k) H[k] = Hn[k];
    ho_hinv(H, Hi);
    int ov = overlap_all(c, H, Hi, R, 0);
    if (mode == HO_MODE_PROPOSE) {
        set_dec(out, !ov, ov ? HO_REJ_OVERLAP : HO_ACCEPT, ov ? 0.0 : boltz_f);
        free(bak);
        return;
    }
    if (ov) { /* alkane_box_resize(.., reset=1): restore the saved cell and chains exactly */
        for (size_t k = 0; k < n3; ++k) R[k] = bak[k];
        for (int k = 0; k < 9; ++k) H[k] = bak[n3 + k];
        set_dec(out, 0, HO_REJ_OVERLAP, 0.0);
    } else {
        set_dec(out, 1, HO_ACCEPT, boltz_f);
    }
    free(bak);
}

/* box_shear_step (MCNS.py:135-207) and box_stretch_step (MCNS.py:209-252) + MCNS.py:360-368 */
static void move_shape(const ho_cfg *c, double *H, double *R, const ho_proposal *p, int mode,
                       ho_decision *out)
{
    double dH[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 };
    if (p->type == HO_SHEAR) {
        int k = p->idx0;
        int o0 = (k == 0) ? 1 : 0;
        int o1 = (k == 2) ? 1 : 2;
        double v1[3] = { H[3 * o0], H[3 * o0 + 1], H[3 * o0 + 2] };
        double v2[3] = { H[3 * o1], H[3 * o1 + 1], H[3 * o1 + 2] };
        double n1 = sqrt(dot3_blas(v1, v1));               /* MCNS.py:161 */
        v1[0] /= n1; v1[1] /= n1; v1[2] /= n1;
        double pr = dot3_blas(v1, v2);                /* MCNS.py:162 */
        v2[0] -= v1[0] * pr; v2[1] -= v1[1] * pr; v2[2] -= v1[2] * pr;
        double n2 = sqrt(dot3_blas(v2, v2));               /* MCNS.py:163 */
        v2[0] /= n2; v2[1] /= n2; v2[2] /= n2;
        for (int a = 0; a < 3; ++a) {                   /* MCNS.py:182-184 */
            double nw = H[3 * k + a] + (p->p[0] * v1[a] + p->p[1] * v2[a]);
            dH[3 * k + a] = nw - H[3 * k + a];
        }
    } else {
        int i = p->idx0, j = p->idx1;
        double e1 = hd_exp(p->p[0]), e2 = hd_exp(-p->p[0]);
        for (int a = 0; a < 3; ++a) {                   /* MCNS.py:231-234 */
            dH[3 * i + a] = H[3 * i + a] * e1 - H[3 * i + a];
            dH[3 * j + a] = H[3 * j + a] * e2 - H[3 * j + a];
        }
    }
    ho_change_box(c, H, R, dH);                         /* MCNS.py:186,238 */
    int reason = HO_ACCEPT;
    if (ho_aspect_fails(H, c->min_ar)) reason = HO_REJ_ASPECT;              /* MCNS.py:194,243 */
    else if (ho_angle_fails(H, c->cos_min_ang)) reason = HO_REJ_ANGLE;      /* MCNS.py:196,245 */
    else if (ho_check_overlap(c, H, R)) reason = HO_REJ_OVERLAP;            /* MCNS.py:198,247 */
    double boltz = reason == HO_ACCEPT ? 1.0 : 0.0;
    set_dec(out, reason == HO_ACCEPT, reason, boltz);
    if (mode == HO_MODE_PROPOSE || reason == HO_ACCEPT) return;
    double ndH[9];
    for (int k = 0; k < 9; ++k) ndH[k] = -1.0 * dH[k];  /* MCNS.py:366-368: NOT an exact restore */
    ho_change_box(c, H, R, ndH);
}

void ho_apply_proposal(const ho_cfg *c, double *H, double *R, const ho_proposal *p,
                       double volume_limit, int mode, ho_decision *out)
{
    switch (p->type) {
    case HO_VOL: move_volume(c, H, R, p, volume_limit, mode, out); break;
    case HO_TRANS:
    case HO_ROT:
    case HO_DIH: move_chain(c, H, R, p, mode, out); break;
    default: move_shape(c, H, R, p, mode, out); break;
    }
}

void ho_replay(const ho_cfg *c, double *H, double *R, const ho_proposal *p, int nmoves,
               double volume_limit, int mode, ho_decision *out)
{
    for (int m = 0; m < nmoves; ++m) ho_apply_proposal(c, H, R, p + m, volume_limit, mode, out + m);
}

/* MC_run, NesSa/MCNS.py:276-392 */
double ho_mc_run(const ho_cfg *c, double *H, double *R, int sweeps, double volume_limit,
                 uint64_t seed, uint32_t gid, uint64_t *ctr, uint64_t attempted[6],
                 uint64_t accepted[6])
{
    int mps = ho_moves_per_sweep(c->nbeads, c->nchains);
    uint64_t m = *ctr;
    for (int isw = 0; isw < sweeps; ++isw)
        for (int im = 0; im < mps; ++im, ++m) {
            ho_proposal p;
            ho_decision d;
            ho_gen_proposal(c, seed, gid, m, &p);
            ho_apply_proposal(c, H, R, &p, volume_limit, HO_MODE_FUSED, &d);
            attempted[p.type] += 1;
            accepted[p.type] += (uint64_t)d.accepted;
        }
    *ctr = m;
    return ho_volume(H);
}

void ho_mc_run_many(const ho_cfg *c, double *H, double *R, const int32_t *ids, int n, int sweeps,
                    const double *volume_limit, uint64_t seed, const uint32_t *gids,
                    uint64_t *ctrs, uint64_t *attempted, uint64_t *accepted, double *vols,
                    int nthreads)
{
    const size_t n3 = (size_t)3 * c->nchains * c->nbeads;
    (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads > 0 ? nthreads : omp_get_max_threads())
#endif
    for (int k = 0; k < n; ++k) {
        int w = ids[k];
        vols[k] = ho_mc_run(c, H + 9 * (size_t)w, R + n3 * (size_t)w, sweeps, volume_limit[k], seed,
                            gids[k], ctrs + w, attempted + 6 * (size_t)k, accepted + 6 * (size_t)k);
    }
}

/* ---------------------------------------------------------------------------------------------
 * Chain growth [A9].  hs_alkane grows chains by configurational-bias MC; the reference only
 * needs ANY valid non-overlapping start because it is randomised by initial_walk sweeps
 * afterwards (MCNS.py:628-644, mpihans.py:164-165), so this is plain random insertion with
 * fixed bond length / bond angle and uniform dihedrals, rejected on the first overlap.
 * Draws: Philox blocks keyed by (seed; attempt, ichain, walker id), stream 1.
 * ------------------------------------------------------------------------------------------- */
int ho_grow_chain(const ho_cfg *c, const double *H, double *R, int ichain, uint64_t seed,
                  uint32_t gid, uint32_t attempt)
{
    const int nb = c->nbeads;
    double u[HO_MAX_BEADS + 6] = { 0.0 };
    int nblk = (nb + 3 + 1) / 2;
    for (int b = 0; b < nblk; ++b) draw_block(seed, attempt, (uint32_t)ichain, gid, (uint32_t)b, HO_STREAM_GROW, u + 2 * b);
    double x[3 * HO_MAX_BEADS];
    double f[3] = { u[0], u[1], u[2] };
    vecmat(f, H, x);
    if (nb > 1) {
        double z = 2.0 * u[3] - 1.0;
        double st2 = 1.0 - z * z;
        double st = sqrt(st2 > 0.0 ? st2 : 0.0);
        double s, co;
        hd_sincos(HD_TWO_PI * u[4], &s, &co);
        x[3] = x[0] + c->bondlength * (st * co);
        x[4] = x[1] + c->bondlength * (st * s);
        x[5] = x[2] + c->bondlength * z;
    }
    for (int k = 2; k < nb; ++k) {
        double b[3] = { x[3 * (k - 1)] - x[3 * (k - 2)], x[3 * (k - 1) + 1] - x[3 * (k - 2) + 1],
                        x[3 * (k - 1) + 2] - x[3 * (k - 2) + 2] };
        double bn = sqrt(dot3(b, b));
        double uh[3] = { b[0] / bn, b[1] / bn, b[2] / bn };
        int ax = 0;
        if (fabs(uh[1]) < fabs(uh[ax])) ax = 1;
        if (fabs(uh[2]) < fabs(uh[ax])) ax = 2;
        double e[3] = { 0.0, 0.0, 0.0 }, e1[3], e2[3];
        e[ax] = 1.0;
        cross3(uh, e, e1);
        double n1 = sqrt(dot3(e1, e1));
        e1[0] /= n1; e1[1] /= n1; e1[2] /= n1;
        cross3(uh, e1, e2);
        double s, co;
        hd_sincos(HD_TWO_PI * u[5 + (k - 2)], &s, &co);
        for (int a = 0; a < 3; ++a) {
            double nd = -c->cos_bondangle * uh[a] + c->sin_bondangle * (co * e1[a] + s * e2[a]);
            x[3 * k + a] = x[3 * (k - 1) + a] + c->bondlength * nd;
        }
    }
    double Hi[9];
    ho_hinv(H, Hi);
    for (int k = 0; k < nb; ++k) {
        for (int jc = 0; jc < ichain; ++jc)
            for (int jb = 0; jb < nb; ++jb)
                if (ho_pair_r2(H, Hi, x + 3 * k, R_AT(R, c, jc, jb)) < 1.0) return 1;
        for (int j = 0; j < k - c->nexclude; ++j)
            if (ho_pair_r2(H, Hi, x + 3 * j, x + 3 * k) < 1.0) return 1;
    }
    double *dst = R_AT(R, c, ichain, 0);
    for (int k = 0; k < 3 * nb; ++k) dst[k] = x[k];
    return 0;
}

/* populate_boxes for one walker (MCNS.py:634-644): retry each chain until it fits.
 * Returns the number of attempts used, or -1 if a chain did not fit in max_attempts. */
int ho_grow_walker(const ho_cfg *c, const double *H, double *R, uint64_t seed, uint32_t gid,
                   int max_attempts)
{
    int total = 0;
    for (int ic = 0; ic < c->nchains; ++ic) {
        int ok = 0;
        for (int a = 0; a < max_attempts; ++a) {
            ++total;
            if (ho_grow_chain(c, H, R, ic, seed, gid, (uint32_t)a) == 0) { ok = 1; break; }
        }
        if (!ok) return -1;
    }
    return total;
}
