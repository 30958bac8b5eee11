// hans_kernels.cu -- sm_100a kernels + C ABI of libhans_b200.so (see include/hans_b200.h).
//
// Execution model: one WALKER per thread group of T threads (T = 32: one warp per walker, four
// walkers per CTA; T = 64/128/256: one CTA per walker).  The walker's beads live in shared memory
// as float64 structure-of-arrays for the whole walk; the cell H and its inverse are kept in
// shared memory and copied to registers around the pair loops.  All lanes of a group compute
// the (cheap, scalar) proposal and acceptance arithmetic redundantly so control flow is uniform
// across the group and no broadcast is needed; the O(N) / O(N^2) pair loops are split across the
// lanes and combined with a group-wide vote that doubles as the early exit.
//
// Compile: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -lineinfo  (see
// __graft_entry__.build()).  -fmad=false is REQUIRED for bit-exact parity with the CPU checker.
#include "hans_device.cuh"

#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>

namespace hb {

// ---------------------------------------------------------------------------------------------
// group primitives
// ---------------------------------------------------------------------------------------------
template <int T> __device__ __forceinline__ void gsync()
{
    if (T == 32) __syncwarp(); else __syncthreads();
}
template <int T> __device__ __forceinline__ bool gany(bool p)
{
    if (T == 32) return __any_sync(FULL, p);
    return __syncthreads_or(p ? 1 : 0) != 0;
}

// one proposal as it sits in the shared-memory ring (same 64-byte layout as hans_proposal)
struct __align__(16) PropRec {
    int type, ichain, idx0, idx1;
    double p[4];
    double u_acc;
    double pad;
};

// per-walker working set in shared memory
struct WCtx {
    int N, nb, nc, nex, npc;
    int lane;
    double *X, *Y, *Z;    // current beads (float64 master)
    double *TX, *TY, *TZ; // trial beads (box moves)
    double *CN;           // trial chain, [3*b + k]
    double *SH;           // [0..8] H, [9..17] H^-1, [18..26] trial H, [27..35] trial H^-1
    float *FX, *FY, *FZ;  // float32 shadow of the current beads: wrapped fractional coordinates
    float *GX, *GY, *GZ;  // the same for the trial beads
    float *FCN;           // the same for the trial chain
    float *SF;            // filter parameters: [0..7] for H, [8..15] for the trial H (see filter_params)
    PropRec *ring;        // T proposals generated lane-parallel
    const uchar2 *itab;   // intra-chain pairs (j, b), b - j > nexclude
    unsigned long long pairs;
};

__host__ __device__ inline int intra_pairs_per_chain(int nb, int nex)
{
    const int m = nb - nex - 1;
    return m > 0 ? m * (m + 1) / 2 : 0;
}

// bytes of shared memory one walker group of T threads needs
__host__ __device__ inline size_t walker_smem_bytes(int N, int nb, int T)
{
    size_t b = sizeof(double) * ((size_t)6 * N + (size_t)3 * nb + 36);
    b += sizeof(float) * ((size_t)6 * N + (size_t)3 * nb + 16);
    b = (b + 15) & ~(size_t)15;
    b += sizeof(PropRec) * (size_t)T;
    return b;
}

// beads [b0, nb) of the trial chain CN against every bead of every other chain
template <int T>
__device__ bool chain_inter_overlap(WCtx &w, int ichain, int b0, const double *H, const double *Hi)
{
    const int nb = w.nb, c0 = ichain * nb, M = w.N - nb;
    bool ov = false;
    for (int base = 0; base < M; base += T) {
        const int jj = base + w.lane;
        if (jj < M) {
            const int j = jj < c0 ? jj : jj + nb;
            const double xj = w.X[j], yj = w.Y[j], zj = w.Z[j];
            for (int b = b0; b < nb; ++b) {
                const double r2 = pair_r2(H, Hi, xj - w.CN[3 * b], yj - w.CN[3 * b + 1], zj - w.CN[3 * b + 2]);
                ov |= (r2 < 1.0);
            }
            w.pairs += (unsigned)(nb - b0);
        }
        if (gany<T>(ov)) return true;
    }
    return false;
}

// intra-chain pairs (j, b) of the trial chain CN with b >= b0
template <int T>
__device__ bool chain_intra_overlap(WCtx &w, int b0, const double *H, const double *Hi)
{
    bool ov = false;
    for (int e = w.lane; e < w.npc; e += T) {
        const uchar2 jb = w.itab[e];
        if ((int)jb.y >= b0) {
            const double *a = w.CN + 3 * jb.x, *b = w.CN + 3 * jb.y;
            ov |= (pair_r2(H, Hi, b[0] - a[0], b[1] - a[1], b[2] - a[2]) < 1.0);
            w.pairs += 1;
        }
    }
    return gany<T>(ov);
}

// all inter-chain pairs (and optionally the intra-chain ones) of the beads in (X,Y,Z).
// Bead-level round robin: lane owns bead i (registers) and walks the offsets d = 1..N/2 with
// j = (i + d) mod N, so a warp reads consecutive shared-memory words (no bank conflicts, no index
// division) and every unordered pair is visited exactly once (for even N the antipodal offset
// d = N/2 is taken by the lower half only).  Same-chain pairs are skipped unless with_intra and
// their separation along the chain exceeds nexclude.
template <int T>
__device__ bool all_overlap(WCtx &w, const double *X, const double *Y, const double *Z, const double *H,
                            const double *Hi, bool with_intra)
{
    const int nb = w.nb, N = w.N, half = N >> 1;
    const bool even = (N & 1) == 0;
    // rows per pass and offset groups: when the group has more lanes than beads, split the offsets
    const int RB = T < N ? T : N;
    const int DG = T < N ? 1 : T / N;
    const int li = w.lane % RB, dg = w.lane / RB;
    const bool lane_on = dg < DG;
    bool ov = false;
    constexpr int U = 4; // offsets per lane between votes
    for (int i0 = 0; i0 < N; i0 += RB) {
        const int i = i0 + li;
        const bool row_on = lane_on && i < N;
        const int ii = row_on ? i : 0;
        const double xi = X[ii], yi = Y[ii], zi = Z[ii];
        const int c0 = (ii / nb) * nb, c1 = c0 + nb;
        for (int d0 = 1; d0 <= half; d0 += DG * U) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int d = d0 + u * DG + dg;
                if (row_on && d <= half && !(even && d == half && i >= half)) {
                    int j = i + d;
                    if (j >= N) j -= N;
                    const bool same = (j >= c0) && (j < c1);
                    const int sep = j > i ? j - i : i - j;
                    if (!same || (with_intra && sep > w.nex)) {
                        ov |= (pair_r2(H, Hi, X[j] - xi, Y[j] - yi, Z[j] - zi) < 1.0);
                        w.pairs += 1;
                    }
                }
            }
            if (gany<T>(ov)) return true;
        }
    }
    return false;
}

// chains follow a cell change rigidly, anchored on their first bead: (X) under Hi_old -> (TX)
// under H_new.  src and dst may not alias.
template <int T>
__device__ void follow_cell(const WCtx &w, const double *sx, const double *sy, const double *sz, double *dx,
                            double *dy, double *dz, const double *Hi_old, const double *H_new)
{
    for (int i = w.lane; i < w.N; i += T) {
        const int a = (i / w.nb) * w.nb;
        const double ax = sx[a], ay = sy[a], az = sz[a];
        double f0, f1, f2, n0, n1, n2;
        vecmat(ax, ay, az, Hi_old, f0, f1, f2);
        vecmat(f0, f1, f2, H_new, n0, n1, n2);
        dx[i] = sx[i] + (n0 - ax);
        dy[i] = sy[i] + (n1 - ay);
        dz[i] = sz[i] + (n2 - az);
    }
}

// ---------------------------------------------------------------------------------------------
// float32 pre-filter of the pair test.
//
// The decision "r^2 < 1" is always the float64 one (pair_r2, the arithmetic of the CPU checker).
// To reach it cheaply, every bead also has a float32 shadow: its fractional coordinates wrapped
// to [-0.5, 0.5].  For a pair, ds = s_j - s_i is wrapped again and r^2 = ds.G.ds is evaluated
// with the metric tensor G = H H^T in float32 (21 operations instead of 32 float64 ones).  If the
// result is outside the band [1 - eps, 1 + eps] the float64 test is guaranteed to agree and is
// skipped; inside the band (probability ~1e-4) the lane evaluates the float64 test.
//
// eps bounds |r2_f32 - r2_f64| for pairs near contact: inputs carry <= 1.5 * 2^-24 per fractional
// component (two stored roundings + one subtraction), which moves r^2 by <= 2 |G ds| . |d(ds)|
// <= 9 Lmax 2^-24 |d|; the nine float32 operations and the rounding of G add <= ~10 * 2^-24 times
// the sum of |terms| <= (sum_i L_i / w_i)^2 |d|^2 (L_i cell vector lengths, w_i perpendicular widths).
// eps = 2^-24 (16 Lmax + 16 q^2 + 64) * 1.01 carries a further safety factor; tests/ checks that no
// float32 classification ever contradicts the float64 one.  The filter is switched off (every
// pair goes to float64) when a perpendicular width is <= 2.1, where float32 and float64 could wrap a
// pair to different images that both lie within reach of sigma.
// SF layout: g00, g11, g22, 2 g01, 2 g02, 2 g12, lo = 1 - eps, hi = 1 + eps.
// ---------------------------------------------------------------------------------------------
__device__ __noinline__ void filter_params(const double *H, const double *Hi, float *out8, bool writer)
{
    const double g00 = dot3(H, H), g11 = dot3(H + 3, H + 3), g22 = dot3(H + 6, H + 6);
    const double g01 = dot3(H, H + 3), g02 = dot3(H, H + 6), g12 = dot3(H + 3, H + 6);
    const double L0 = sqrt(g00), L1 = sqrt(g11), L2 = sqrt(g22);
    const double b0 = sqrt((Hi[0] * Hi[0] + Hi[3] * Hi[3]) + Hi[6] * Hi[6]);
    const double b1 = sqrt((Hi[1] * Hi[1] + Hi[4] * Hi[4]) + Hi[7] * Hi[7]);
    const double b2 = sqrt((Hi[2] * Hi[2] + Hi[5] * Hi[5]) + Hi[8] * Hi[8]);
    const double Lmax = fmax(L0, fmax(L1, L2));
    const double bmax = fmax(b0, fmax(b1, b2)); // 1 / smallest perpendicular width
    const double q = (L0 * b0 + L1 * b1) + L2 * b2;
    double eps = 5.9604644775390625e-8 * ((16.0 * Lmax + 16.0 * q * q) + 64.0) * 1.01;
    float lo = (float)(1.0 - eps), hi = (float)(1.0 + eps);
    if (!(bmax * 2.1 < 1.0) || !(eps < 0.05)) { lo = -1.0f; hi = 3.0e38f; } // filter off
    if (writer) {
        out8[0] = (float)g00; out8[1] = (float)g11; out8[2] = (float)g22;
        out8[3] = (float)(2.0 * g01); out8[4] = (float)(2.0 * g02); out8[5] = (float)(2.0 * g12);
        out8[6] = lo; out8[7] = hi;
    }
}

struct Filt { float g00, g11, g22, g01, g02, g12, lo, hi; };
__device__ __forceinline__ Filt load_filt(const float *sf)
{
    Filt f;
    f.g00 = sf[0]; f.g11 = sf[1]; f.g22 = sf[2]; f.g01 = sf[3]; f.g02 = sf[4]; f.g12 = sf[5]; f.lo = sf[6]; f.hi = sf[7];
    return f;
}

__device__ __forceinline__ float rint32(float x)
{
    const float magic = 12582912.0f; // 1.5 * 2^23
    const float t = __fadd_rn(x, magic);
    return __fsub_rn(t, magic);
}

// float32 r^2 of a fractional difference (not yet wrapped)
__device__ __forceinline__ float r2_f32(const Filt &f, float dx, float dy, float dz)
{
    dx = __fsub_rn(dx, rint32(dx));
    dy = __fsub_rn(dy, rint32(dy));
    dz = __fsub_rn(dz, rint32(dz));
    const float t0 = __fmaf_rn(f.g02, dz, __fmaf_rn(f.g01, dy, __fmul_rn(f.g00, dx)));
    const float t1 = __fmaf_rn(f.g12, dz, __fmul_rn(f.g11, dy));
    const float t2 = __fmul_rn(f.g22, dz);
    return __fmaf_rn(dx, t0, __fmaf_rn(dy, t1, __fmul_rn(dz, t2)));
}

// wrapped fractional coordinates of a Cartesian point, rounded to float32
__device__ __forceinline__ void shadow_of(double x, double y, double z, const double *Hi, float &fx, float &fy, float &fz)
{
    double s0, s1, s2;
    vecmat(x, y, z, Hi, s0, s1, s2);
    fx = (float)(s0 - hd_rint(s0));
    fy = (float)(s1 - hd_rint(s1));
    fz = (float)(s2 - hd_rint(s2));
}

// float64 decisions for the rare pairs inside the band (kept out of line: cold)
__device__ __noinline__ bool exact_chain_pair(const WCtx &w, int j, int b)
{
    return pair_r2(w.SH, w.SH + 9, w.X[j] - w.CN[3 * b], w.Y[j] - w.CN[3 * b + 1], w.Z[j] - w.CN[3 * b + 2]) < 1.0;
}
__device__ __noinline__ bool exact_intra_pair(const WCtx &w, int j, int b)
{
    const double *p = w.CN + 3 * j, *q = w.CN + 3 * b;
    return pair_r2(w.SH, w.SH + 9, q[0] - p[0], q[1] - p[1], q[2] - p[2]) < 1.0;
}
__device__ __noinline__ bool exact_trial_pair(const WCtx &w, int i, int j)
{
    return pair_r2(w.SH + 18, w.SH + 27, w.TX[j] - w.TX[i], w.TY[j] - w.TY[i], w.TZ[j] - w.TZ[i]) < 1.0;
}

// beads [b0, nb) of the trial chain against every bead of every other chain (filtered)
template <int T>
__device__ __forceinline__ bool chain_inter_overlap_f(WCtx &w, int ichain, int b0)
{
    const Filt f = load_filt(w.SF);
    const int nb = w.nb, c0 = ichain * nb, M = w.N - nb;
    bool ov = false;
    for (int base = 0; base < M; base += T) {
        const int jj = base + w.lane;
        if (jj < M) {
            const int j = jj < c0 ? jj : jj + nb;
            const float fx = w.FX[j], fy = w.FY[j], fz = w.FZ[j];
            for (int b = b0; b < nb; ++b) {
                const float r2 = r2_f32(f, fx - w.FCN[3 * b], fy - w.FCN[3 * b + 1], fz - w.FCN[3 * b + 2]);
                if (r2 < f.lo) ov = true;
                else if (!(r2 > f.hi)) ov |= exact_chain_pair(w, j, b);
            }
            w.pairs += (unsigned)(nb - b0);
        }
        if (gany<T>(ov)) return true;
    }
    return false;
}

// intra-chain pairs (j, b) of the trial chain with b >= b0 (filtered)
template <int T>
__device__ bool chain_intra_overlap_f(WCtx &w, int b0)
{
    const Filt f = load_filt(w.SF);
    bool ov = false;
    for (int e = w.lane; e < w.npc; e += T) {
        const uchar2 jb = w.itab[e];
        if ((int)jb.y >= b0) {
            const float *p = w.FCN + 3 * jb.x, *q = w.FCN + 3 * jb.y;
            const float r2 = r2_f32(f, q[0] - p[0], q[1] - p[1], q[2] - p[2]);
            if (r2 < f.lo) ov = true;
            else if (!(r2 > f.hi)) ov |= exact_intra_pair(w, jb.x, jb.y);
            w.pairs += 1;
        }
    }
    return gany<T>(ov);
}

// all pairs of the TRIAL beads (filtered); same bead-level round robin as all_overlap
template <int T>
__device__ __noinline__ bool trial_overlap_f(WCtx &w, bool with_intra)
{
    const Filt f = load_filt(w.SF + 8);
    const int nb = w.nb, N = w.N, half = N >> 1;
    const bool even = (N & 1) == 0;
    const int RB = T < N ? T : N;
    const int DG = T < N ? 1 : T / N;
    const int li = w.lane % RB, dg = w.lane / RB;
    const bool lane_on = dg < DG;
    bool ov = false;
    constexpr int U = 4;
    for (int i0 = 0; i0 < N; i0 += RB) {
        const int i = i0 + li;
        const bool row_on = lane_on && i < N;
        const int ii = row_on ? i : 0;
        const float xi = w.GX[ii], yi = w.GY[ii], zi = w.GZ[ii];
        const int c0 = (ii / nb) * nb, c1 = c0 + nb;
        for (int d0 = 1; d0 <= half; d0 += DG * U) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int d = d0 + u * DG + dg;
                if (row_on && d <= half && !(even && d == half && i >= half)) {
                    int j = i + d;
                    if (j >= N) j -= N;
                    const bool same = (j >= c0) && (j < c1);
                    const int sep = j > i ? j - i : i - j;
                    if (!same || (with_intra && sep > w.nex)) {
                        const float r2 = r2_f32(f, w.GX[j] - xi, w.GY[j] - yi, w.GZ[j] - zi);
                        if (r2 < f.lo) ov = true;
                        else if (!(r2 > f.hi)) ov |= exact_trial_pair(w, i, j);
                        w.pairs += 1;
                    }
                }
            }
            if (gany<T>(ov)) return true;
        }
    }
    return false;
}

// chains follow a cell change rigidly (first bead keeps its fractional coordinates): current beads
// under Hi_old -> trial beads under H_new, plus their float32 shadows under Hi_new
template <int T>
__device__ void follow_to_trial(const WCtx &w, const double *Hi_old, const double *H_new, const double *Hi_new)
{
    for (int i = w.lane; i < w.N; i += T) {
        const int a = (i / w.nb) * w.nb;
        const double ax = w.X[a], ay = w.Y[a], az = w.Z[a];
        double f0, f1, f2, n0, n1, n2;
        vecmat(ax, ay, az, Hi_old, f0, f1, f2);
        vecmat(f0, f1, f2, H_new, n0, n1, n2);
        const double x = w.X[i] + (n0 - ax), y = w.Y[i] + (n1 - ay), z = w.Z[i] + (n2 - az);
        w.TX[i] = x; w.TY[i] = y; w.TZ[i] = z;
        shadow_of(x, y, z, Hi_new, w.GX[i], w.GY[i], w.GZ[i]);
    }
}

// the reverse direction for a rejected shear / stretch: trial beads under Hi_old -> current beads
// under H_new (MCNS.py:366-368 reverts by applying -dH, not by restoring)
template <int T>
__device__ void follow_from_trial(const WCtx &w, const double *Hi_old, const double *H_new, const double *Hi_new)
{
    for (int i = w.lane; i < w.N; i += T) {
        const int a = (i / w.nb) * w.nb;
        const double ax = w.TX[a], ay = w.TY[a], az = w.TZ[a];
        double f0, f1, f2, n0, n1, n2;
        vecmat(ax, ay, az, Hi_old, f0, f1, f2);
        vecmat(f0, f1, f2, H_new, n0, n1, n2);
        const double x = w.TX[i] + (n0 - ax), y = w.TY[i] + (n1 - ay), z = w.TZ[i] + (n2 - az);
        w.X[i] = x; w.Y[i] = y; w.Z[i] = z;
        shadow_of(x, y, z, Hi_new, w.FX[i], w.FY[i], w.FZ[i]);
    }
}

template <int T> __device__ void commit_trial(WCtx &w)
{
    for (int i = w.lane; i < w.N; i += T) {
        w.X[i] = w.TX[i]; w.Y[i] = w.TY[i]; w.Z[i] = w.TZ[i];
        w.FX[i] = w.GX[i]; w.FY[i] = w.GY[i]; w.FZ[i] = w.GZ[i];
    }
    if (w.lane < 18) w.SH[w.lane] = w.SH[18 + w.lane];
    if (w.lane < 8) w.SF[w.lane] = w.SF[8 + w.lane];
    gsync<T>();
}

struct Dec { int accepted, reason; double boltz; };
__device__ __forceinline__ Dec mkdec(int a, int r, double b) { Dec d; d.accepted = a; d.reason = r; d.boltz = b; return d; }

// translate / rotate / dihedral: alk.alkane_translate_chain, alkane_rotate_chain,
// alkane_bond_rotate (MCNS.py:323,328,333) + accept/restore (MCNS.py:371-380)
template <int T>
__device__ __forceinline__ Dec move_chain(WCtx &w, const PropRec &p, int mode)
{
    const int nb = w.nb, c0 = p.ichain * nb;
    int b0 = 0;
    bool intra = false;
    if (p.type == HANS_TRANS) {
        for (int b = w.lane; b < nb; b += T) {
            const double x = w.X[c0 + b] + p.p[0], y = w.Y[c0 + b] + p.p[1], z = w.Z[c0 + b] + p.p[2];
            w.CN[3 * b] = x; w.CN[3 * b + 1] = y; w.CN[3 * b + 2] = z;
            shadow_of(x, y, z, w.SH + 9, w.FCN[3 * b], w.FCN[3 * b + 1], w.FCN[3 * b + 2]);
        }
    } else if (p.type == HANS_ROT) {
        double com[3] = {0.0, 0.0, 0.0};
        for (int b = 0; b < nb; ++b) { com[0] += w.X[c0 + b]; com[1] += w.Y[c0 + b]; com[2] += w.Z[c0 + b]; }
        const double dn = (double)nb;
        com[0] /= dn; com[1] /= dn; com[2] /= dn;
        for (int b = w.lane; b < nb; b += T) {
            const double v[3] = {w.X[c0 + b] - com[0], w.Y[c0 + b] - com[1], w.Z[c0 + b] - com[2]};
            double o[3];
            quat_rot(p.p, v, o);
            const double x = o[0] + com[0], y = o[1] + com[1], z = o[2] + com[2];
            w.CN[3 * b] = x; w.CN[3 * b + 1] = y; w.CN[3 * b + 2] = z;
            shadow_of(x, y, z, w.SH + 9, w.FCN[3 * b], w.FCN[3 * b + 1], w.FCN[3 * b + 2]);
        }
    } else {
        const int ia = p.idx0;
        if (nb < 4 || ia < 3 || ia > nb - 1) return mkdec(0, HANS_REJ_INVALID, 0.0);
        const bool flip = p.idx1 != 0;
        const int iv = c0 + (flip ? nb - 1 - (ia - 1) : ia - 1);
        const int iw = c0 + (flip ? nb - 1 - (ia - 2) : ia - 2);
        const double piv[3] = {w.X[iv], w.Y[iv], w.Z[iv]};
        double ax[3] = {piv[0] - w.X[iw], piv[1] - w.Y[iw], piv[2] - w.Z[iw]};
        const double n = sqrt(dot3(ax, ax));
        ax[0] /= n; ax[1] /= n; ax[2] /= n;
        double s, co;
        hd_sincos(0.5 * p.p[0], &s, &co);
        const double q[4] = {co, s * ax[0], s * ax[1], s * ax[2]};
        for (int b = w.lane; b < nb; b += T) {
            const int src = c0 + (flip ? nb - 1 - b : b);
            double r[3] = {w.X[src], w.Y[src], w.Z[src]};
            if (b >= ia) {
                const double v[3] = {r[0] - piv[0], r[1] - piv[1], r[2] - piv[2]};
                double o[3];
                quat_rot(q, v, o);
                r[0] = o[0] + piv[0]; r[1] = o[1] + piv[1]; r[2] = o[2] + piv[2];
            }
            w.CN[3 * b] = r[0]; w.CN[3 * b + 1] = r[1]; w.CN[3 * b + 2] = r[2];
            shadow_of(r[0], r[1], r[2], w.SH + 9, w.FCN[3 * b], w.FCN[3 * b + 1], w.FCN[3 * b + 2]);
        }
        b0 = ia;
        intra = true;
    }
    gsync<T>();
    bool ov = chain_inter_overlap_f<T>(w, p.ichain, b0);
    if (!ov && intra) ov = chain_intra_overlap_f<T>(w, b0);
    const double boltz = ov ? 0.0 : 1.0;
    const bool keep = (mode == HANS_MODE_PROPOSE) || (p.u_acc < boltz); // MCNS.py:372
    if (keep) {
        for (int b = w.lane; b < nb; b += T) {
            w.X[c0 + b] = w.CN[3 * b]; w.Y[c0 + b] = w.CN[3 * b + 1]; w.Z[c0 + b] = w.CN[3 * b + 2];
            w.FX[c0 + b] = w.FCN[3 * b]; w.FY[c0 + b] = w.FCN[3 * b + 1]; w.FZ[c0 + b] = w.FCN[3 * b + 2];
        }
    }
    gsync<T>();
    return mkdec(ov ? 0 : 1, ov ? HANS_REJ_OVERLAP : HANS_ACCEPT, boltz);
}

// alk.alkane_box_resize(pressure, ibox, 0) + acceptance (MCNS.py:318,349-357).  The trial lives
// in (TX,TY,TZ), so a rejection restores the saved cell and chains exactly for free.
template <int T>
__device__ __noinline__ Dec move_volume(WCtx &w, const hans_cfg &c, const PropRec &p, double volume_limit, int mode)
{
    double H[9], Hi[9], Hn[9], Hin[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) { H[k] = w.SH[k]; Hi[k] = w.SH[9 + k]; }
    const double vol = fabs(det3(H));
    const double vnew_t = vol + p.p[0];
    if (!(vnew_t > 0.0)) return mkdec(0, HANS_REJ_INVALID, 0.0);
    const double ratio = vnew_t / vol;
    const double s = hd_cbrt(ratio);
#pragma unroll
    for (int k = 0; k < 9; ++k) Hn[k] = H[k] * s;
    const double new_volume = fabs(det3(Hn)); // MCNS.py:350
    const double boltz_f = hd_exp(-c.pressure * p.p[0] + (double)c.nchains * hd_log(ratio)); // MCNS.py:266
    if (mode == HANS_MODE_FUSED) {
        // conjunction of MCNS.py:351 evaluated cheapest-first; the O(N^2) overlap test is last
        if (!((new_volume - volume_limit) < 2.2204460492503131e-16)) return mkdec(0, HANS_REJ_CEILING, boltz_f);
        if (!(p.u_acc < boltz_f)) return mkdec(0, HANS_REJ_BOLTZ, boltz_f);
    }
    hinv(Hn, Hin);
    follow_to_trial<T>(w, Hi, Hn, Hin);
    filter_params(Hn, Hin, w.SF + 8, w.lane == 0);
    if (w.lane == 0) {
#pragma unroll
        for (int k = 0; k < 9; ++k) { w.SH[18 + k] = Hn[k]; w.SH[27 + k] = Hin[k]; }
    }
    gsync<T>();
    const bool ov = trial_overlap_f<T>(w, false);
    if (mode == HANS_MODE_PROPOSE || !ov) commit_trial<T>(w);
    if (ov) return mkdec(0, HANS_REJ_OVERLAP, 0.0);
    return mkdec(1, HANS_ACCEPT, boltz_f);
}

// box_shear_step (MCNS.py:135-207), box_stretch_step (MCNS.py:209-252), accept / revert by -dH
// (MCNS.py:360-368)
template <int T>
__device__ __noinline__ Dec move_shape(WCtx &w, const hans_cfg &c, const PropRec &p, int mode)
{
    double H[9], Hi[9], Hn[9], Hin[9], dH[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) { H[k] = w.SH[k]; Hi[k] = w.SH[9 + k]; dH[k] = 0.0; }
    if (p.type == HANS_SHEAR) {
        const int k = p.idx0;
        const int o0 = (k == 0) ? 1 : 0;
        const int o1 = (k == 2) ? 1 : 2;
        double v1[3] = {H[3 * o0], H[3 * o0 + 1], H[3 * o0 + 2]};
        double v2[3] = {H[3 * o1], H[3 * o1 + 1], H[3 * o1 + 2]};
        const double n1 = sqrt(dot3(v1, v1));
        v1[0] /= n1; v1[1] /= n1; v1[2] /= n1;
        const double pr = dot3(v1, v2);
        v2[0] -= v1[0] * pr; v2[1] -= v1[1] * pr; v2[2] -= v1[2] * pr;
        const double n2 = sqrt(dot3(v2, v2));
        v2[0] /= n2; v2[1] /= n2; v2[2] /= n2;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double nw = H[3 * k + a] + (p.p[0] * v1[a] + p.p[1] * v2[a]);
            dH[3 * k + a] = nw - H[3 * k + a];
        }
    } else {
        const int i = p.idx0, j = p.idx1;
        const double e1 = hd_exp(p.p[0]), e2 = hd_exp(-p.p[0]);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            dH[3 * i + a] = H[3 * i + a] * e1 - H[3 * i + a];
            dH[3 * j + a] = H[3 * j + a] * e2 - H[3 * j + a];
        }
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) Hn[k] = H[k] + dH[k];
    hinv(Hn, Hin);
    follow_to_trial<T>(w, Hi, Hn, Hin); // alk.alkane_change_box, MCNS.py:186,238
    filter_params(Hn, Hin, w.SF + 8, w.lane == 0);
    if (w.lane == 0) {
#pragma unroll
        for (int k = 0; k < 9; ++k) { w.SH[18 + k] = Hn[k]; w.SH[27 + k] = Hin[k]; }
    }
    gsync<T>();
    int reason = HANS_ACCEPT;
    if (min_aspect_ratio(Hn) < c.min_ar) reason = HANS_REJ_ASPECT;         // MCNS.py:194,243
    else if (max_abs_cos(Hn) > c.cos_min_ang) reason = HANS_REJ_ANGLE;     // MCNS.py:196,245
    else if (trial_overlap_f<T>(w, true)) reason = HANS_REJ_OVERLAP;       // MCNS.py:198,247
    if (mode == HANS_MODE_PROPOSE || reason == HANS_ACCEPT) {
        commit_trial<T>(w);
    } else {
        // MCNS.py:366-368: the reference reverts by applying -dH through alkane_change_box, which
        // is not an exact restore; reproduce that arithmetic from the trial state.
        double H2[9], H2i[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) H2[k] = Hn[k] + (-1.0 * dH[k]);
        hinv(H2, H2i);
        follow_from_trial<T>(w, Hin, H2, H2i);
        filter_params(H2, H2i, w.SF, w.lane == 0);
        if (w.lane == 0) {
#pragma unroll
            for (int k = 0; k < 9; ++k) { w.SH[k] = H2[k]; w.SH[9 + k] = H2i[k]; }
        }
        gsync<T>();
    }
    return mkdec(reason == HANS_ACCEPT ? 1 : 0, reason, reason == HANS_ACCEPT ? 1.0 : 0.0);
}

template <int T>
__device__ __forceinline__ Dec apply_move(WCtx &w, const hans_cfg &c, const PropRec &p, double volume_limit, int mode)
{
    if (p.type == HANS_VOL) return move_volume<T>(w, c, p, volume_limit, mode);
    if (p.type <= HANS_DIH) return move_chain<T>(w, p, mode);
    return move_shape<T>(w, c, p, mode);
}

// ---------------------------------------------------------------------------------------------
// shared-memory carve-up and state load/store
// ---------------------------------------------------------------------------------------------
template <int T>
__device__ void setup_ctx(WCtx &w, const hans_cfg &c, double *smem, int groups_per_cta)
{
    w.nb = c.nbeads; w.nc = c.nchains; w.N = c.nbeads * c.nchains; w.nex = c.nexclude;
    w.npc = intra_pairs_per_chain(w.nb, w.nex);
    w.lane = threadIdx.x % T;
    const int g = threadIdx.x / T;
    const size_t per = walker_smem_bytes(w.N, w.nb, T);
    char *base = reinterpret_cast<char *>(smem) + per * g;
    w.X = reinterpret_cast<double *>(base); w.Y = w.X + w.N; w.Z = w.Y + w.N;
    w.TX = w.Z + w.N; w.TY = w.TX + w.N; w.TZ = w.TY + w.N;
    w.CN = w.TZ + w.N;
    w.SH = w.CN + 3 * w.nb;
    w.FX = reinterpret_cast<float *>(w.SH + 36); w.FY = w.FX + w.N; w.FZ = w.FY + w.N;
    w.GX = w.FZ + w.N; w.GY = w.GX + w.N; w.GZ = w.GY + w.N;
    w.FCN = w.GZ + w.N;
    w.SF = w.FCN + 3 * w.nb;
    w.ring = reinterpret_cast<PropRec *>(base + per - sizeof(PropRec) * (size_t)T);
    uchar2 *tab = reinterpret_cast<uchar2 *>(reinterpret_cast<char *>(smem) + per * groups_per_cta);
    w.itab = tab;
    w.pairs = 0;
    // intra-chain pair table, ordered by b then j (one per CTA)
    for (int e = threadIdx.x; e < w.npc; e += blockDim.x) {
        int b = w.nex + 1, acc = 0;
        while (acc + (b - w.nex) <= e) { acc += b - w.nex; ++b; }
        tab[e] = make_uchar2((unsigned char)(e - acc), (unsigned char)b);
    }
    __syncthreads();
}

template <int T>
__device__ void load_walker(WCtx &w, const double *gH, const double *gR, int wi)
{
    const double *r = gR + (size_t)wi * w.N * 3;
    double H[9], Hi[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) H[k] = gH[(size_t)wi * 9 + k];
    hinv(H, Hi);
    for (int i = w.lane; i < w.N; i += T) {
        const double x = r[3 * i], y = r[3 * i + 1], z = r[3 * i + 2];
        w.X[i] = x; w.Y[i] = y; w.Z[i] = z;
        shadow_of(x, y, z, Hi, w.FX[i], w.FY[i], w.FZ[i]);
    }
    filter_params(H, Hi, w.SF, w.lane == 0);
    if (w.lane == 0) {
#pragma unroll
        for (int k = 0; k < 9; ++k) { w.SH[k] = H[k]; w.SH[9 + k] = Hi[k]; }
    }
    gsync<T>();
}

template <int T>
__device__ void store_walker(const WCtx &w, double *gH, double *gR, int wi)
{
    double *r = gR + (size_t)wi * w.N * 3;
    for (int i = w.lane; i < w.N; i += T) { r[3 * i] = w.X[i]; r[3 * i + 1] = w.Y[i]; r[3 * i + 2] = w.Z[i]; }
    if (w.lane < 9) gH[(size_t)wi * 9 + w.lane] = w.SH[w.lane];
}

template <int T> __device__ void flush_pairs(WCtx &w, unsigned long long *g_pairs)
{
    if (g_pairs == nullptr) return;
    unsigned long long v = w.pairs;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(FULL, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(g_pairs, v);
    w.pairs = 0;
}

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
struct WalkArgs {
    hans_cfg cfg;
    double *H, *R;
    unsigned long long *ctr;
    const int32_t *ids;
    int n_active, nmoves;
    const double *d_limit;
    double limit;
    unsigned long long seed;
    uint32_t gid_base;
    unsigned long long *attempted, *accepted;
    double *vol;
    unsigned long long *pair_tests;
    // replay
    const hans_proposal *props;
    hans_decision *out;
    int mode;
};

template <int T> __host__ __device__ constexpr int cta_threads() { return T == 32 ? 128 : T; }

// MC_run (MCNS.py:276-392) for a batch of walkers; REPLAY = proposals come from memory.
// Moves are processed in batches of T: first every lane produces the proposal of one move of the
// batch (Philox + the transcendental part of the proposal run once per move instead of once per
// lane), then the batch is applied move by move by the whole group.
template <int T, bool REPLAY>
__global__ void __launch_bounds__(cta_threads<T>()) walk_kernel(const WalkArgs a)
{
    extern __shared__ double smem[];
    constexpr int GPC = cta_threads<T>() / T;
    WCtx w;
    setup_ctx<T>(w, a.cfg, smem, GPC);
    const int g = threadIdx.x / T;
    const double limit = a.d_limit ? *a.d_limit : a.limit;
    for (int k0 = blockIdx.x * GPC; k0 < a.n_active; k0 += gridDim.x * GPC) {
        const int k = k0 + g;
        if (k < a.n_active) {
            const int wi = a.ids ? a.ids[k] : k;
            load_walker<T>(w, a.H, a.R, wi);
            const unsigned long long ctr0 = REPLAY ? 0ull : a.ctr[wi];
            unsigned att[6] = {0, 0, 0, 0, 0, 0}, acc[6] = {0, 0, 0, 0, 0, 0}; // this lane's share
            for (int m0 = 0; m0 < a.nmoves; m0 += T) {
                const int nbatch = a.nmoves - m0 < T ? a.nmoves - m0 : T;
                int my_type = -1;
                if (w.lane < nbatch) {
                    PropRec *slot = w.ring + w.lane;
                    if (REPLAY) {
                        const uint4 *q = reinterpret_cast<const uint4 *>(a.props + (size_t)k * a.nmoves + m0 + w.lane);
                        uint4 *d = reinterpret_cast<uint4 *>(slot);
                        d[0] = q[0]; d[1] = q[1]; d[2] = q[2]; d[3] = q[3];
                        my_type = slot->type;
                    } else {
                        const Prop p = gen_proposal(a.cfg, a.seed, a.gid_base + (uint32_t)wi,
                                                    ctr0 + (unsigned long long)(m0 + w.lane));
                        slot->type = p.type; slot->ichain = p.ichain; slot->idx0 = p.idx0; slot->idx1 = p.idx1;
                        slot->p[0] = p.p[0]; slot->p[1] = p.p[1]; slot->p[2] = p.p[2]; slot->p[3] = p.p[3];
                        slot->u_acc = p.u_acc;
                        my_type = p.type;
                    }
                }
                gsync<T>();
                int my_acc = 0, my_reason = 0;
                double my_boltz = 0.0;
                for (int q = 0; q < nbatch; ++q) {
                    const Dec d = apply_move<T>(w, a.cfg, w.ring[q], limit, a.mode);
                    if (w.lane == q) { my_acc = d.accepted; my_reason = d.reason; my_boltz = d.boltz; }
                }
                if (my_type >= 0) {
#pragma unroll
                    for (int t = 0; t < 6; ++t) { att[t] += (my_type == t); acc[t] += (my_type == t && my_acc); }
                    if (REPLAY) {
                        hans_decision *o = a.out + (size_t)k * a.nmoves + m0 + w.lane;
                        o->accepted = my_acc; o->reason = my_reason; o->boltz = my_boltz;
                    }
                }
                gsync<T>(); // the ring is rewritten by the next batch
            }
            store_walker<T>(w, a.H, a.R, wi);
            if (a.attempted) {
                // reduce the per-lane counters over the group
#pragma unroll
                for (int t = 0; t < 6; ++t) {
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        att[t] += __shfl_down_sync(FULL, att[t], o);
                        acc[t] += __shfl_down_sync(FULL, acc[t], o);
                    }
                    if ((threadIdx.x & 31) == 0) {
                        atomicAdd(a.attempted + (size_t)k * 6 + t, (unsigned long long)att[t]);
                        atomicAdd(a.accepted + (size_t)k * 6 + t, (unsigned long long)acc[t]);
                    }
                }
            }
            if (w.lane == 0) {
                if (!REPLAY) a.ctr[wi] = ctr0 + (unsigned long long)a.nmoves;
                if (a.vol) a.vol[wi] = fabs(det3(w.SH));
            }
            flush_pairs<T>(w, a.pair_tests);
            gsync<T>();
        }
        if (T != 32) __syncthreads();
    }
}

// alk.alkane_check_chain_overlap (MCNS.py:198,247,543) for a batch of walkers
template <int T>
__global__ void __launch_bounds__(cta_threads<T>())
overlap_kernel(const hans_cfg cfg, const double *gH, const double *gR, const int32_t *ids, int n, int inter_only,
               int32_t *out)
{
    extern __shared__ double smem[];
    constexpr int GPC = cta_threads<T>() / T;
    WCtx w;
    setup_ctx<T>(w, cfg, smem, GPC);
    const int g = threadIdx.x / T;
    for (int k0 = blockIdx.x * GPC; k0 < n; k0 += gridDim.x * GPC) {
        const int k = k0 + g;
        if (k < n) {
            const int wi = ids ? ids[k] : k;
            load_walker<T>(w, gH, gR, wi);
            double H[9], Hi[9];
#pragma unroll
            for (int q = 0; q < 9; ++q) { H[q] = w.SH[q]; Hi[q] = w.SH[9 + q]; }
            const bool ov = all_overlap<T>(w, w.X, w.Y, w.Z, H, Hi, !inter_only);
            if (w.lane == 0) out[k] = ov ? 1 : 0;
            gsync<T>();
        }
        if (T != 32) __syncthreads();
    }
}

// alk.alkane_change_box(ibox, dH) (MCNS.py:186,238,368) for a batch of walkers
template <int T>
__global__ void __launch_bounds__(cta_threads<T>())
change_box_kernel(const hans_cfg cfg, double *gH, double *gR, const int32_t *ids, int n, const double *gdH)
{
    extern __shared__ double smem[];
    constexpr int GPC = cta_threads<T>() / T;
    WCtx w;
    setup_ctx<T>(w, cfg, smem, GPC);
    const int g = threadIdx.x / T;
    for (int k0 = blockIdx.x * GPC; k0 < n; k0 += gridDim.x * GPC) {
        const int k = k0 + g;
        if (k < n) {
            const int wi = ids ? ids[k] : k;
            load_walker<T>(w, gH, gR, wi);
            double Hi[9], Hn[9];
#pragma unroll
            for (int q = 0; q < 9; ++q) { Hi[q] = w.SH[9 + q]; Hn[q] = w.SH[q] + gdH[(size_t)k * 9 + q]; }
            follow_cell<T>(w, w.X, w.Y, w.Z, w.TX, w.TY, w.TZ, Hi, Hn);
            gsync<T>();
            if (w.lane == 0) {
#pragma unroll
                for (int q = 0; q < 9; ++q) w.SH[18 + q] = Hn[q];
            }
            gsync<T>();
            commit_trial<T>(w);
            store_walker<T>(w, gH, gR, wi);
            gsync<T>();
        }
        if (T != 32) __syncthreads();
    }
}

// create_initial_configs / populate_boxes (MCNS.py:628-644); alk.alkane_grow_chain (MCNS.py:642).
// Random insertion with fixed bond length / angle and uniform dihedrals; a chain is retried with
// the next attempt counter until it overlaps nothing already placed.
template <int T>
__global__ void __launch_bounds__(cta_threads<T>())
grow_kernel(const hans_cfg cfg, const double *gH, double *gR, const int32_t *ids, int n, unsigned long long seed,
            uint32_t gid_base, const uint32_t *gid_override, int chain_begin, int chain_end, uint32_t attempt_begin,
            int max_attempts, int32_t *status)
{
    extern __shared__ double smem[];
    constexpr int GPC = cta_threads<T>() / T;
    WCtx w;
    setup_ctx<T>(w, cfg, smem, GPC);
    const int g = threadIdx.x / T;
    const int nb = w.nb;
    for (int k0 = blockIdx.x * GPC; k0 < n; k0 += gridDim.x * GPC) {
        const int k = k0 + g;
        if (k < n) {
            const int wi = ids ? ids[k] : k;
            const uint32_t gid = gid_override ? gid_override[k] : gid_base + (uint32_t)wi;
            load_walker<T>(w, gH, gR, wi);
            double H[9], Hi[9];
#pragma unroll
            for (int q = 0; q < 9; ++q) { H[q] = w.SH[q]; Hi[q] = w.SH[9 + q]; }
            const int cend = chain_end < 0 ? w.nc : chain_end;
            int total = 0;
            bool failed = false;
            for (int ic = chain_begin; ic < cend && !failed; ++ic) {
                bool placed = false;
                for (int at = 0; at < max_attempts && !placed; ++at) {
                    ++total;
                    const uint32_t attempt = attempt_begin + (uint32_t)at;
                    // every lane builds the same trial chain; uniforms u[0..nb+2]
                    double ua, ub, uc, ud, ue, uf;
                    draw_block(seed, attempt, (uint32_t)ic, gid, 0u, STREAM_GROW, ua, ub);
                    draw_block(seed, attempt, (uint32_t)ic, gid, 1u, STREAM_GROW, uc, ud);
                    draw_block(seed, attempt, (uint32_t)ic, gid, 2u, STREAM_GROW, ue, uf);
                    double x0, y0, z0;
                    vecmat(ua, ub, uc, H, x0, y0, z0);
                    double pp[3] = {x0, y0, z0}, pc[3] = {x0, y0, z0}; // previous-previous, previous
                    if (w.lane == 0) { w.CN[0] = x0; w.CN[1] = y0; w.CN[2] = z0; }
                    if (nb > 1) {
                        const double z = 2.0 * ud - 1.0;
                        const double st2 = 1.0 - z * z;
                        const double st = sqrt(st2 > 0.0 ? st2 : 0.0);
                        double s, co;
                        hd_sincos(HD_TWO_PI * ue, &s, &co);
                        pc[0] = x0 + cfg.bondlength * (st * co);
                        pc[1] = y0 + cfg.bondlength * (st * s);
                        pc[2] = z0 + cfg.bondlength * z;
                        if (w.lane == 0) { w.CN[3] = pc[0]; w.CN[4] = pc[1]; w.CN[5] = pc[2]; }
                    }
                    double ucur_a = ue, ucur_b = uf; // block 2 holds u[4], u[5]
                    uint32_t cur_blk = 2u;
                    for (int kb = 2; kb < nb; ++kb) {
                        const int ui = 5 + (kb - 2); // index of the uniform for this bead
                        const uint32_t blk = (uint32_t)(ui >> 1);
                        if (blk != cur_blk) { draw_block(seed, attempt, (uint32_t)ic, gid, blk, STREAM_GROW, ucur_a, ucur_b); cur_blk = blk; }
                        const double u = (ui & 1) ? ucur_b : ucur_a;
                        const double b[3] = {pc[0] - pp[0], pc[1] - pp[1], pc[2] - pp[2]};
                        const double bn = sqrt(dot3(b, b));
                        const double uh[3] = {b[0] / bn, b[1] / bn, b[2] / bn};
                        int ax = 0;
                        if (fabs(uh[1]) < fabs(uh[ax])) ax = 1;
                        if (fabs(uh[2]) < fabs(uh[ax])) ax = 2;
                        double e[3] = {0.0, 0.0, 0.0}, e1[3], e2[3];
                        e[ax] = 1.0;
                        cross3(uh, e, e1);
                        const double n1 = sqrt(dot3(e1, e1));
                        e1[0] /= n1; e1[1] /= n1; e1[2] /= n1;
                        cross3(uh, e1, e2);
                        double s, co;
                        hd_sincos(HD_TWO_PI * u, &s, &co);
                        double nw[3];
#pragma unroll
                        for (int q = 0; q < 3; ++q) {
                            const double nd = -cfg.cos_bondangle * uh[q] + cfg.sin_bondangle * (co * e1[q] + s * e2[q]);
                            nw[q] = pc[q] + cfg.bondlength * nd;
                        }
                        if (w.lane == 0) { w.CN[3 * kb] = nw[0]; w.CN[3 * kb + 1] = nw[1]; w.CN[3 * kb + 2] = nw[2]; }
                        pp[0] = pc[0]; pp[1] = pc[1]; pp[2] = pc[2];
                        pc[0] = nw[0]; pc[1] = nw[1]; pc[2] = nw[2];
                    }
                    gsync<T>();
                    // against everything already placed (chains < ic), then intra-chain
                    bool ov = false;
                    const int M = ic * nb;
                    for (int base = 0; base < M; base += T) {
                        const int j = base + w.lane;
                        if (j < M) {
                            const double xj = w.X[j], yj = w.Y[j], zj = w.Z[j];
                            for (int b = 0; b < nb; ++b)
                                ov |= (pair_r2(H, Hi, xj - w.CN[3 * b], yj - w.CN[3 * b + 1], zj - w.CN[3 * b + 2]) < 1.0);
                        }
                        if (gany<T>(ov)) break;
                    }
                    ov = gany<T>(ov);
                    if (!ov) ov = chain_intra_overlap<T>(w, 0, H, Hi);
                    if (!ov) {
                        for (int b = w.lane; b < nb; b += T) {
                            w.X[ic * nb + b] = w.CN[3 * b]; w.Y[ic * nb + b] = w.CN[3 * b + 1]; w.Z[ic * nb + b] = w.CN[3 * b + 2];
                        }
                        placed = true;
                    }
                    gsync<T>();
                }
                if (!placed) failed = true;
            }
            store_walker<T>(w, const_cast<double *>(gH), gR, wi);
            if (w.lane == 0 && status) status[k] = failed ? -1 : total;
            gsync<T>();
        }
        if (T != 32) __syncthreads();
    }
}

// alk.box_compute_volume / box_min_aspect_ratio / min_angle (MCNS.py:97-133) for n cells
__global__ void metrics_kernel(const double *gH, int n, double *vol, double *mar, double *mcos)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double H[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) H[k] = gH[(size_t)i * 9 + k];
    if (vol) vol[i] = fabs(det3(H));
    if (mar) mar[i] = min_aspect_ratio(H);
    if (mcos) mcos[i] = max_abs_cos(H);
}

// local MAXLOC (mpihans.py:211-212): ties -> lowest index
__global__ void argmax_kernel(const double *vol, int n, double *out2)
{
    __shared__ double sv[32];
    __shared__ int si[32];
    double best = -1.7976931348623157e308;
    int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double v = vol[i];
        if (v > best || (v == best && i < bi)) { best = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_down_sync(FULL, best, o);
        const int oi = __shfl_down_sync(FULL, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int nw = (blockDim.x + 31) >> 5;
        best = threadIdx.x < nw ? sv[threadIdx.x] : -1.7976931348623157e308;
        bi = threadIdx.x < nw ? si[threadIdx.x] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_down_sync(FULL, best, o);
            const int oi = __shfl_down_sync(FULL, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (threadIdx.x == 0) { out2[0] = best; out2[1] = (double)bi; }
    }
}

// clone (mpihans.py:226-237)
__global__ void clone_kernel(int N, const double *sH, const double *sR, const double *svol, int src,
                             const int32_t *d_src, double *dH, double *dR, double *dvol, int dst,
                             const int32_t *d_dst, const int32_t *d_enable)
{
    if (d_enable && *d_enable == 0) return;
    const int s = d_src ? *d_src : src;
    const int d = d_dst ? *d_dst : dst;
    const double *r = sR + (size_t)s * N * 3;
    double *o = dR + (size_t)d * N * 3;
    if (r == o) return;
    for (int i = threadIdx.x; i < 3 * N; i += blockDim.x) o[i] = r[i];
    if (threadIdx.x < 9) dH[(size_t)d * 9 + threadIdx.x] = sH[(size_t)s * 9 + threadIdx.x];
    if (threadIdx.x == 0 && dvol && svol) dvol[d] = svol[s];
}

// ---------------------------------------------------------------------------------------------
// nested-sampling iteration helpers (mpihans.py:210-245), single CTA each
// ---------------------------------------------------------------------------------------------
constexpr uint32_t STREAM_NS = 2u;

__device__ __forceinline__ unsigned long long ns_draw(unsigned long long seed, long long it, uint32_t c2, uint32_t blk)
{
    uint32_t x[4];
    philox4x32_10((uint32_t)it, (uint32_t)((unsigned long long)it >> 32), c2, blk | (STREAM_NS << 16), (uint32_t)seed,
                  (uint32_t)(seed >> 32), x);
    return ((unsigned long long)x[0] << 32) | x[1];
}

__device__ void block_argmax(const double *vol, int n, double &best, int &bi)
{
    __shared__ double sv[32];
    __shared__ int si[32];
    best = -1.7976931348623157e308;
    bi = 0x7fffffff;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double v = vol[i];
        if (v > best || (v == best && i < bi)) { best = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_down_sync(FULL, best, o);
        const int oi = __shfl_down_sync(FULL, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int nw = (blockDim.x + 31) >> 5;
        best = threadIdx.x < nw ? sv[threadIdx.x] : -1.7976931348623157e308;
        bi = threadIdx.x < nw ? si[threadIdx.x] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_down_sync(FULL, best, o);
            const int oi = __shfl_down_sync(FULL, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (threadIdx.x == 0) { sv[0] = best; si[0] = bi; }
    }
    __syncthreads();
    best = sv[0];
    bi = si[0];
    __syncthreads();
}

// local MAXLOC (mpihans.py:211-212) and staging of the clone source (mpihans.py:226-229)
__global__ void ns_pack_kernel(const hans_ns_args a)
{
    double best; int bi;
    block_argmax(a.d_vol, a.nlocal, best, bi);
    const long long it = *a.d_iter;
    const unsigned long long K = (unsigned long long)a.nlocal * (unsigned long long)a.world;
    const unsigned long long g = ns_draw(a.seed, it, 0u, 0u) % K; // replaces randint, mpihans.py:220
    const int src_rank = (int)(g / (unsigned long long)a.nlocal), src = (int)(g % (unsigned long long)a.nlocal);
    if (threadIdx.x == 0) { a.d_rec[0] = best; a.d_rec[1] = (double)bi; }
    if (src_rank == a.rank) {
        const int n3 = 3 * a.N;
        if (threadIdx.x == 0) a.d_rec[2] = a.d_vol[src];
        if (threadIdx.x < 9) a.d_rec[3 + threadIdx.x] = a.d_H[(size_t)src * 9 + threadIdx.x];
        const double *r = a.d_R + (size_t)src * n3;
        for (int i = threadIdx.x; i < n3; i += blockDim.x) a.d_rec[12 + i] = r[i];
    }
}

// global MAXLOC (mpihans.py:215), ceiling, log, trajectory staging, clone (mpihans.py:232-237),
// active walkers (mpihans.py:236-239)
__global__ void ns_resolve_kernel(const hans_ns_args a)
{
    __shared__ double s_vmax;
    __shared__ int s_rank, s_idx;
    const int rec = 12 + 3 * a.N, n3 = 3 * a.N;
    const long long it = *a.d_iter;
    if (threadIdx.x == 0) {
        double best = -1.7976931348623157e308; int br = 0, bi = 0;
        for (int r = 0; r < a.world; ++r) {
            const double v = a.d_gathered[(size_t)r * rec];
            if (v > best) { best = v; br = r; bi = (int)a.d_gathered[(size_t)r * rec + 1]; }
        }
        s_vmax = best; s_rank = br; s_idx = bi;
        *a.d_limit = best;
        if (a.d_log && a.log_cap > 0) {
            double *l = a.d_log + 3 * (size_t)(it % a.log_cap);
            l[0] = best; l[1] = (double)br; l[2] = (double)bi;
        }
    }
    __syncthreads();
    const int mrank = s_rank, midx = s_idx;
    const unsigned long long K = (unsigned long long)a.nlocal * (unsigned long long)a.world;
    const unsigned long long g = ns_draw(a.seed, it, 0u, 0u) % K;
    const int src_rank = (int)(g / (unsigned long long)a.nlocal);
    if (mrank == a.rank) {
        double *dr = a.d_R + (size_t)midx * n3;
        if (a.traj_interval > 0 && a.d_traj && (it % a.traj_interval) == 0) {
            const int slot = (int)((it / a.traj_interval) % a.traj_slots);
            double *t = a.d_traj + (size_t)slot * rec;
            if (threadIdx.x == 0) { t[0] = s_vmax; t[1] = (double)midx; t[2] = a.d_vol[midx]; a.d_traj_iter[slot] = it; }
            if (threadIdx.x < 9) t[3 + threadIdx.x] = a.d_H[(size_t)midx * 9 + threadIdx.x];
            for (int i = threadIdx.x; i < n3; i += blockDim.x) t[12 + i] = dr[i];
            __syncthreads();
        }
        const double *sr = a.d_gathered + (size_t)src_rank * rec;
        if (threadIdx.x == 0) a.d_vol[midx] = sr[2];
        if (threadIdx.x < 9) a.d_H[(size_t)midx * 9 + threadIdx.x] = sr[3 + threadIdx.x];
        for (int i = threadIdx.x; i < n3; i += blockDim.x) dr[i] = sr[12 + i];
    }
    const int nv = a.nlocal / a.nw_per_vrank;
    for (int v = threadIdx.x; v < nv; v += blockDim.x) {
        int id;
        if (mrank == a.rank && midx / a.nw_per_vrank == v) id = midx;                       // mpihans.py:236
        else id = v * a.nw_per_vrank + (int)(ns_draw(a.seed, it, (uint32_t)(a.rank * nv + v), 1u) % (unsigned)a.nw_per_vrank); // :239
        a.d_ids[v] = id;
    }
    __syncthreads();
    if (threadIdx.x == 0) *a.d_iter = it + 1;
}

// FP32 FFMA issue-rate probe (roofline denominator)
__global__ void ffma_kernel(int iters, float *sink)
{
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f;
    float a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
    const float m = 0.999f, c = 1e-3f;
    for (int i = 0; i < iters; ++i) {
        a0 = __fmaf_rn(a0, m, c); a1 = __fmaf_rn(a1, m, c); a2 = __fmaf_rn(a2, m, c); a3 = __fmaf_rn(a3, m, c);
        a4 = __fmaf_rn(a4, m, c); a5 = __fmaf_rn(a5, m, c); a6 = __fmaf_rn(a6, m, c); a7 = __fmaf_rn(a7, m, c);
    }
    const float s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 123456.789f) sink[0] = s;
}

} // namespace hb

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
using namespace hb;

static thread_local std::string g_err;
static int fail(int code, const std::string &msg) { g_err = msg; return code; }
static int cuda_fail(cudaError_t e, const char *what)
{
    return fail(HANS_ECUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cuda_fail(e_, #call); } while (0)

static int check_cfg(const hans_cfg *c)
{
    if (!c) return fail(HANS_EINVAL, "cfg is NULL");
    if (c->nchains < 1 || c->nbeads < 1) return fail(HANS_EINVAL, "nchains and nbeads must be >= 1");
    if (c->nbeads > HANS_MAX_BEADS) return fail(HANS_ELIMIT, "nbeads exceeds HANS_MAX_BEADS");
    if (c->nchains > 32767) return fail(HANS_ELIMIT, "nchains too large");
    if (c->nexclude < 0) return fail(HANS_EINVAL, "nexclude must be >= 0");
    return HANS_OK;
}

static size_t cta_smem_bytes(const hans_cfg *c, int T)
{
    const int gpc = (T == 32) ? 4 : 1;
    const int N = c->nchains * c->nbeads;
    size_t b = walker_smem_bytes(N, c->nbeads, T) * gpc;
    b += (size_t)intra_pairs_per_chain(c->nbeads, c->nexclude) * sizeof(uchar2);
    return (b + 15) & ~(size_t)15;
}

static int sm_count()
{
    static int n = 0;
    if (!n) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 148;
    }
    return n;
}

template <typename K> static int prep_kernel(K kernel, size_t smem)
{
    if (smem > 227 * 1024) return fail(HANS_ELIMIT, "walker does not fit in shared memory (cell-list path not built yet)");
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return HANS_OK;
}

static int grid_for(int n, int T)
{
    const int gpc = (T == 32) ? 4 : 1;
    const int ctas = (n + gpc - 1) / gpc;
    const int cap = sm_count() * 16;
    return ctas < cap ? (ctas > 0 ? ctas : 1) : cap;
}

static int pick_T(const hans_cfg *c, int T)
{
    if (T == 0) T = hans_default_threads_per_walker(c);
    if (T != 32 && T != 64 && T != 128 && T != 256) return -1;
    return T;
}

#define DISPATCH_T(T, ...)                                    \
    switch (T) {                                              \
    case 32: { constexpr int TT = 32; __VA_ARGS__; } break;   \
    case 64: { constexpr int TT = 64; __VA_ARGS__; } break;   \
    case 128: { constexpr int TT = 128; __VA_ARGS__; } break; \
    default: { constexpr int TT = 256; __VA_ARGS__; } break;  \
    }

extern "C" {

int hans_abi_version(void) { return HANS_ABI_VERSION; }
const char *hans_last_error(void) { return g_err.c_str(); }

int hans_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int64_t hans_walker_smem_bytes(const hans_cfg *cfg)
{
    if (check_cfg(cfg) != HANS_OK) return -1;
    return (int64_t)walker_smem_bytes(cfg->nchains * cfg->nbeads, cfg->nbeads, hans_default_threads_per_walker(cfg));
}

int hans_moves_per_sweep(int nbeads, int nchains)
{
    if (nbeads == 1) return nchains + 7;
    if (nbeads <= 3) return 2 * nchains + 7;
    return (nbeads - 1) * nchains + 7;
}

int hans_default_threads_per_walker(const hans_cfg *cfg)
{
    if (!cfg) return 32;
    const int N = cfg->nchains * cfg->nbeads;
    if (N <= 128) return 32;
    if (N <= 256) return 64;
    if (N <= 512) return 128;
    return 256;
}

int hans_mc_run_batch(const hans_cfg *cfg, double *d_H, double *d_R, uint64_t *d_ctr, const int32_t *d_ids,
                      int n_active, int sweeps, const double *d_volume_limit, double volume_limit, uint64_t seed,
                      uint32_t gid_base, uint64_t *d_attempted, uint64_t *d_accepted, double *d_vol,
                      uint64_t *d_pair_tests, int threads_per_walker, void *stream)
{
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if (!d_H || !d_R || !d_ctr) return fail(HANS_EINVAL, "d_H, d_R and d_ctr are required");
    if (n_active < 0 || sweeps < 0) return fail(HANS_EINVAL, "negative n_active / sweeps");
    if ((d_attempted == nullptr) != (d_accepted == nullptr)) return fail(HANS_EINVAL, "d_attempted and d_accepted go together");
    if (n_active == 0) return HANS_OK;
    const int T = pick_T(cfg, threads_per_walker);
    if (T < 0) return fail(HANS_EINVAL, "threads_per_walker must be 0, 32, 64, 128 or 256");
    WalkArgs a;
    memset(&a, 0, sizeof a);
    a.cfg = *cfg; a.H = d_H; a.R = d_R; a.ctr = (unsigned long long *)d_ctr; a.ids = d_ids; a.n_active = n_active;
    a.nmoves = sweeps * hans_moves_per_sweep(cfg->nbeads, cfg->nchains);
    a.d_limit = d_volume_limit; a.limit = volume_limit; a.seed = seed; a.gid_base = gid_base;
    a.attempted = (unsigned long long *)d_attempted; a.accepted = (unsigned long long *)d_accepted;
    a.vol = d_vol; a.pair_tests = (unsigned long long *)d_pair_tests; a.mode = HANS_MODE_FUSED;
    const size_t smem = cta_smem_bytes(cfg, T);
    cudaStream_t st = (cudaStream_t)stream;
    DISPATCH_T(T, {
        rc = prep_kernel(walk_kernel<TT, false>, smem);
        if (rc) return rc;
        walk_kernel<TT, false><<<grid_for(n_active, TT), cta_threads<TT>(), smem, st>>>(a);
    });
    CK(cudaGetLastError());
    return HANS_OK;
}

int hans_replay(const hans_cfg *cfg, double *d_H, double *d_R, const int32_t *d_ids, int n_active,
                const hans_proposal *d_props, int nmoves, double volume_limit, int mode, hans_decision *d_out,
                int threads_per_walker, void *stream)
{
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if (!d_H || !d_R || !d_props || !d_out) return fail(HANS_EINVAL, "NULL device pointer");
    if (mode != HANS_MODE_FUSED && mode != HANS_MODE_PROPOSE) return fail(HANS_EINVAL, "bad mode");
    if (n_active < 0 || nmoves < 0) return fail(HANS_EINVAL, "negative count");
    if (n_active == 0 || nmoves == 0) return HANS_OK;
    const int T = pick_T(cfg, threads_per_walker);
    if (T < 0) return fail(HANS_EINVAL, "threads_per_walker must be 0, 32, 64, 128 or 256");
    WalkArgs a;
    memset(&a, 0, sizeof a);
    a.cfg = *cfg; a.H = d_H; a.R = d_R; a.ids = d_ids; a.n_active = n_active; a.nmoves = nmoves;
    a.limit = volume_limit; a.props = d_props; a.out = d_out; a.mode = mode;
    const size_t smem = cta_smem_bytes(cfg, T);
    cudaStream_t st = (cudaStream_t)stream;
    DISPATCH_T(T, {
        rc = prep_kernel(walk_kernel<TT, true>, smem);
        if (rc) return rc;
        walk_kernel<TT, true><<<grid_for(n_active, TT), cta_threads<TT>(), smem, st>>>(a);
    });
    CK(cudaGetLastError());
    return HANS_OK;
}

int hans_check_overlap(const hans_cfg *cfg, const double *d_H, const double *d_R, const int32_t *d_ids, int n,
                       int inter_only, int32_t *d_out, int threads_per_walker, void *stream)
{
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if (!d_H || !d_R || !d_out) return fail(HANS_EINVAL, "NULL device pointer");
    if (n < 0) return fail(HANS_EINVAL, "negative count");
    if (n == 0) return HANS_OK;
    const int T = pick_T(cfg, threads_per_walker);
    if (T < 0) return fail(HANS_EINVAL, "threads_per_walker must be 0, 32, 64, 128 or 256");
    const size_t smem = cta_smem_bytes(cfg, T);
    cudaStream_t st = (cudaStream_t)stream;
    DISPATCH_T(T, {
        rc = prep_kernel(overlap_kernel<TT>, smem);
        if (rc) return rc;
        overlap_kernel<TT><<<grid_for(n, TT), cta_threads<TT>(), smem, st>>>(*cfg, d_H, d_R, d_ids, n, inter_only, d_out);
    });
    CK(cudaGetLastError());
    return HANS_OK;
}

int hans_cell_metrics(const double *d_H, int n, double *d_vol, double *d_min_ar, double *d_max_cos, void *stream)
{
    if (!d_H) return fail(HANS_EINVAL, "d_H is NULL");
    if (n < 0) return fail(HANS_EINVAL, "negative count");
    if (n == 0) return HANS_OK;
    metrics_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(d_H, n, d_vol, d_min_ar, d_max_cos);
    CK(cudaGetLastError());
    return HANS_OK;
}

int hans_change_box(const hans_cfg *cfg, double *d_H, double *d_R, const int32_t *d_ids, int n, const double *d_dH,
                    void *stream)
{
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if (!d_H || !d_R || !d_dH) return fail(HANS_EINVAL, "NULL device pointer");
    if (n < 0) return fail(HANS_EINVAL, "negative count");
    if (n == 0) return HANS_OK;
    const int T = pick_T(cfg, 0);
    const size_t smem = cta_smem_bytes(cfg, T);
    cudaStream_t st = (cudaStream_t)stream;
    DISPATCH_T(T, {
        rc = prep_kernel(change_box_kernel<TT>, smem);
        if (rc) return rc;
        change_box_kernel<TT><<<grid_for(n, TT), cta_threads<TT>(), smem, st>>>(*cfg, d_H, d_R, d_ids, n, d_dH);
    });
    CK(cudaGetLastError());
    return HANS_OK;
}

static int launch_grow(const hans_cfg *cfg, const double *d_H, double *d_R, const int32_t *d_ids, int n, uint64_t seed,
                       uint32_t gid_base, const uint32_t *d_gid, int cb, int ce, uint32_t attempt0, int max_attempts,
                       int32_t *d_status, void *stream)
{
    const int T = pick_T(cfg, 0);
    const size_t smem = cta_smem_bytes(cfg, T);
    cudaStream_t st = (cudaStream_t)stream;
    int rc = HANS_OK;
    DISPATCH_T(T, {
        rc = prep_kernel(grow_kernel<TT>, smem);
        if (rc) return rc;
        grow_kernel<TT><<<grid_for(n, TT), cta_threads<TT>(), smem, st>>>(*cfg, d_H, d_R, d_ids, n, seed, gid_base, d_gid,
                                                                          cb, ce, attempt0, max_attempts, d_status);
    });
    CK(cudaGetLastError());
    return HANS_OK;
}

int hans_grow(const hans_cfg *cfg, const double *d_H, double *d_R, int n, uint64_t seed, uint32_t gid_base,
              int max_attempts, int32_t *d_status, void *stream)
{
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if (!d_H || !d_R) return fail(HANS_EINVAL, "NULL device pointer");
    if (n < 0 || max_attempts < 1) return fail(HANS_EINVAL, "bad count");
    if (n == 0) return HANS_OK;
    return launch_grow(cfg, d_H, d_R, nullptr, n, seed, gid_base, nullptr, 0, -1, 0u, max_attempts, d_status, stream);
}

int hans_grow_chain(const hans_cfg *cfg, const double *d_H, double *d_R, int walker, int ichain, uint64_t seed,
                    uint32_t gid, uint32_t attempt, int32_t *d_ifail, void *stream)
{
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if (!d_H || !d_R || !d_ifail) return fail(HANS_EINVAL, "NULL device pointer");
    if (walker < 0 || ichain < 0 || ichain >= cfg->nchains) return fail(HANS_EINVAL, "bad walker / chain index");
    // one walker, one chain, one attempt: status = 1 (attempts used) on success, -1 on failure
    const size_t off_h = (size_t)walker * 9, off_r = (size_t)walker * cfg->nchains * cfg->nbeads * 3;
    // gid is passed through gid_base with walker index 0 of the shifted views
    return launch_grow(cfg, d_H + off_h, d_R + off_r, nullptr, 1, seed, gid, nullptr, ichain, ichain + 1, attempt, 1,
                       d_ifail, stream);
}

int hans_argmax(const double *d_vol, int n, double *d_out2, void *stream)
{
    if (!d_vol || !d_out2) return fail(HANS_EINVAL, "NULL device pointer");
    if (n < 1) return fail(HANS_EINVAL, "n must be >= 1");
    argmax_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(d_vol, n, d_out2);
    CK(cudaGetLastError());
    return HANS_OK;
}

int hans_clone(const hans_cfg *cfg, const double *d_srcH, const double *d_srcR, const double *d_srcvol, int src,
               const int32_t *d_src_index, double *d_dstH, double *d_dstR, double *d_dstvol, int dst,
               const int32_t *d_dst_index, const int32_t *d_enable, void *stream)
{
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if (!d_srcH || !d_srcR || !d_dstH || !d_dstR) return fail(HANS_EINVAL, "NULL device pointer");
    if (src < 0 || dst < 0) return fail(HANS_EINVAL, "negative walker index");
    clone_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(cfg->nchains * cfg->nbeads, d_srcH, d_srcR, d_srcvol, src, d_src_index,
                                                      d_dstH, d_dstR, d_dstvol, dst, d_dst_index, d_enable);
    CK(cudaGetLastError());
    return HANS_OK;
}

static int check_ns(const hans_ns_args *a)
{
    if (!a) return fail(HANS_EINVAL, "args is NULL");
    if (a->N < 1 || a->nlocal < 1 || a->world < 1 || a->rank < 0 || a->rank >= a->world) return fail(HANS_EINVAL, "bad sizes");
    if (a->nw_per_vrank < 1 || a->nlocal % a->nw_per_vrank) return fail(HANS_EINVAL, "nlocal must be a multiple of nw_per_vrank");
    if (!a->d_H || !a->d_R || !a->d_vol || !a->d_iter || !a->d_rec || !a->d_gathered || !a->d_limit || !a->d_ids)
        return fail(HANS_EINVAL, "NULL device pointer");
    if (a->traj_interval > 0 && a->d_traj && (a->traj_slots < 1 || !a->d_traj_iter)) return fail(HANS_EINVAL, "bad trajectory staging");
    return HANS_OK;
}

int hans_ns_pack(const hans_ns_args *a, void *stream)
{
    int rc = check_ns(a);
    if (rc) return rc;
    ns_pack_kernel<<<1, 512, 0, (cudaStream_t)stream>>>(*a);
    CK(cudaGetLastError());
    return HANS_OK;
}

int hans_ns_resolve(const hans_ns_args *a, void *stream)
{
    int rc = check_ns(a);
    if (rc) return rc;
    ns_resolve_kernel<<<1, 512, 0, (cudaStream_t)stream>>>(*a);
    CK(cudaGetLastError());
    return HANS_OK;
}

int hans_ffma_peak(int iters, double *tflops, void *stream)
{
    if (!tflops || iters < 1) return fail(HANS_EINVAL, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    float *sink = nullptr;
    CK(cudaMalloc(&sink, sizeof(float)));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int blocks = sm_count() * 8, threads = 256;
    ffma_kernel<<<blocks, threads, 0, st>>>(iters, sink); // warm-up
    CK(cudaEventRecord(e0, st));
    ffma_kernel<<<blocks, threads, 0, st>>>(iters, sink);
    CK(cudaEventRecord(e1, st));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    *tflops = 2.0 * 8.0 * (double)iters * blocks * threads / (ms * 1e-3) / 1e12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    return HANS_OK;
}

} // extern "C"
