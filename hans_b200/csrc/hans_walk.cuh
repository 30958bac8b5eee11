// hans_walk.cuh -- the walker kernels (sm_100a): MC_run for a batch of walkers, the overlap
// predicate, alkane_change_box and chain growth.
//
// Execution model: one WALKER per thread group of T threads (T = 32: one warp per walker, four
// walkers per CTA; T = 64/128/256: one CTA per walker).  A walker stays in SHARED MEMORY for its
// whole walk:
//     P   [3N] f64   current beads, x y z interleaved (same layout as the restart file / HBM)
//     TP  [3N] f64   trial beads of a box move (so a rejected volume move restores exactly)
//     CN  [3nb] f64  trial chain of a single-chain move
//     SH  [36] f64   H, H^-1, trial H, trial H^-1
//     F4  [N] float4, G4 [N] float4, C4 [nb] float4   float32 shadows (wrapped fractional
//                    coordinates) of P, TP, CN for the pair-test pre-filter
//     SF  [16] f32   pre-filter parameters for H and for the trial H
//     ring[T] 64 B   proposals of the current batch of T moves
// All shared-memory accesses go through the one `extern __shared__` symbol with integer offsets
// (never through pointers stored in structs), so they compile to LDS/STS and the per-walker
// context is a handful of integers in registers.  The hot path (single-chain moves) is inlined
// and small; box moves (a few percent of the moves, but O(N^2)) are out-of-line functions that
// receive the context by value.
//
// Reference call sites replaced: NesSa/MCNS.py:276-392 (MC_run) and every alk.* call inside it.
#pragma once

#include "hans_device.cuh"

namespace hb {

extern __shared__ double smem[];

// diagnostics: [0] float64 re-tests of a trial chain (band hits), [1] float64 all-pairs passes
__device__ unsigned long long g_diag[4];
#define SMF (reinterpret_cast<float *>(smem))
#define SM4 (reinterpret_cast<float4 *>(smem))

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
template <int T> __host__ __device__ constexpr int cta_threads() { return T == 32 ? 128 : T; }

// one proposal as it sits in the shared-memory ring (same 64-byte layout as hans_proposal)
struct __align__(16) PropRec {
    int type, ichain, idx0, idx1;
    double p[4];
    double u_acc;
    double pad;
};
static_assert(sizeof(PropRec) == 64 && sizeof(hans_proposal) == 64, "proposal record layout");

// per-walker context: offsets into shared memory, kept in registers
struct WCtx {
    int N, nb, nc, nex, npc, lane;
    int D0;   // double index of P; TP = D0+3N, CN = D0+6N, SH = D0+6N+3nb
    int Q0;   // float4 index of F4; G4 = Q0+N, C4 = Q0+2N; SF (floats) = 4*(Q0+2N+nb)
    int ring; // PropRec index (64-byte units from the start of shared memory)
    int itab; // uchar2 index of the intra-chain pair table
    int cfg;  // double index of the CTA's copy of hans_cfg
    int wc;   // speculative windows (hans_lean.cuh): double index of the W trial chains [W][3nb]
    int w4;   //                                      float4 index of their shadows [W][nb]
    int rad;  // float index of the chain bounding radii RAD[nc]
    int lst;  // int index of the near list LST[T] (+ its counter at lst - 1 ... see near_list())
    int cl;   // int index of the per-walker cell list area (cell_smem_bytes), behind the window area / tail
    int cells; // bit 0: every overlap test goes through the per-walker cell list, which the chain moves keep up to date;
               // bit 1: only box moves do (a list of the trial configuration, built per move), and only when the trial
               // configuration is denser than rho_box.  Set per walker by the kernels that carry the area.
    float rho_box;
    unsigned nbmagic, ncmagic; // x / nb = umulhi(x, nbmagic), x / nc = umulhi(x, ncmagic) for x < 2^16
};
#define W_P(w) ((w).D0)
#define W_TP(w) ((w).D0 + 3 * (w).N)
#define W_CN(w) ((w).D0 + 6 * (w).N)
#define W_SH(w) ((w).D0 + 6 * (w).N + 3 * (w).nb)
#define W_F4(w) ((w).Q0)
#define W_G4(w) ((w).Q0 + (w).N)
#define W_C4(w) ((w).Q0 + 2 * (w).N)
#define W_SF(w) (4 * ((w).Q0 + 2 * (w).N + (w).nb))
#define SF_TRIAL 16 /* float offset of the trial cell's filter parameters behind the current cell's */
#define W_RING(w) (reinterpret_cast<PropRec *>(smem) + (w).ring)
#define W_ITAB(w) (reinterpret_cast<const uchar2 *>(smem) + (w).itab)
#define W_CFG(w) (reinterpret_cast<const hans_cfg *>(smem + (w).cfg))

// x / nb and x / nc for 0 <= x < 2^16 without an integer division (the magic overflows for a divisor of 1)
__device__ __forceinline__ int div_nb(const WCtx &w, int x) { return w.nb == 1 ? x : (int)__umulhi((unsigned)x, w.nbmagic); }
__device__ __forceinline__ int div_nc(const WCtx &w, int x) { return w.nc == 1 ? x : (int)__umulhi((unsigned)x, w.ncmagic); }

__host__ __device__ inline int intra_pairs_per_chain(int nb, int nex)
{
    const int m = nb - nex - 1;
    return m > 0 ? m * (m + 1) / 2 : 0;
}

// bytes of shared memory one walker group of T threads needs (multiple of 64)
__host__ __device__ inline size_t walker_smem_bytes(int N, int nb, int T)
{
    size_t b = sizeof(double) * ((size_t)6 * N + (size_t)3 * nb + 36);
    b = (b + 15) & ~(size_t)15;
    b += 16 * ((size_t)2 * N + nb) + 128; // float4 shadows + SF (2 x 16 floats)
    b += (((size_t)4 * (N / nb)) + 15) & ~(size_t)15; // RAD
    b += 16 + 4 * (size_t)T;              // near-list counter + LST
    b = (b + 63) & ~(size_t)63;
    b += 64 * (size_t)T;                 // ring
    return b;
}
__host__ __device__ inline size_t cta_extra_bytes(int nb, int nex)
{
    size_t b = ((sizeof(hans_cfg) + 15) & ~(size_t)15);
    b += (((size_t)intra_pairs_per_chain(nb, nex) * sizeof(uchar2)) + 15) & ~(size_t)15;
    return b;
}

// bytes of the trial chains of a speculative window of W moves (multiple of 64)
__host__ __device__ inline size_t window_smem_bytes(int nb, int W)
{
    return (((size_t)W * nb * 40) + 63) & ~(size_t)63;
}

// Per-walker cell list (the role of hs_alkane's link cells, which the reference bypasses at MCNS.py:603): bins in
// FRACTIONAL space, n_k = floor(w_k / 1.001) per axis (w_k = perpendicular width of the cell along axis k), capped at
// CL_M per axis.  Layout (ints): dims[8] = {n0, n1, n2, nbins} of the current cell and of the trial cell of a box
// move; head2[CL_CAP], next2[N]: the list of the TRIAL configuration of a box move (first bead of each bin, -1 = empty;
// singly linked); head[CL_CAP], next[N]: the list of the current configuration, kept up to date by the chain moves.
// A kernel that uses the list only for box moves (mode 2) allocates the area without the last two arrays.
constexpr int CL_M = 10;
constexpr int CL_CAP = CL_M * CL_M * CL_M;
__host__ __device__ inline size_t cell_smem_bytes(int N, bool box_only = false)
{
    return (4 * ((size_t)8 + (box_only ? 1 : 2) * (CL_CAP + (size_t)N)) + 63) & ~(size_t)63;
}
#define CL_DIMS(w) ((w).cl)
#define CL_HEAD2(w) ((w).cl + 8)
#define CL_NEXT2(w) ((w).cl + 8 + CL_CAP)
#define CL_HEAD(w) ((w).cl + 8 + CL_CAP + (w).N)
#define CL_NEXT(w) ((w).cl + 8 + 2 * CL_CAP + (w).N)
#define SMI (reinterpret_cast<int *>(smem))

// GPC walker groups of T threads per CTA; W > 0 adds the window area behind each walker's ring, `tail` bytes
// (a multiple of 64) behind that (the warp-per-move kernel's scratch and / or the cell list, which starts cl_off bytes
// into the tail)
template <int T, int GPC>
__device__ __forceinline__ WCtx setup_ctx_g(const hans_cfg &c, int W, int tail = 0, int cl_off = 0)
{
    WCtx w;
    w.nb = c.nbeads; w.nc = c.nchains; w.N = c.nbeads * c.nchains; w.nex = c.nexclude;
    w.npc = intra_pairs_per_chain(w.nb, w.nex);
    w.lane = threadIdx.x % T;
    const int g = threadIdx.x / T;
    const int core = (int)walker_smem_bytes(w.N, w.nb, T);
    const int per = core + (int)window_smem_bytes(w.nb, W) + tail;
    const int base = per * g;                                           // bytes
    w.D0 = base / 8;
    const int dbytes = ((8 * (6 * w.N + 3 * w.nb + 36)) + 15) & ~15;
    w.Q0 = (base + dbytes) / 16;
    w.ring = (base + core - 64 * T) / 64;
    {
        const int sf_end = 16 * (w.Q0 + 2 * w.N + w.nb) + 128;          // bytes
        w.rad = sf_end / 4;
        w.lst = (sf_end + ((4 * w.nc + 15) & ~15) + 16) / 4;
    }
    w.nbmagic = 0xFFFFFFFFu / (unsigned)w.nb + 1u;
    w.ncmagic = 0xFFFFFFFFu / (unsigned)w.nc + 1u;
    w.wc = (base + core) / 8;
    w.w4 = (base + core + 24 * W * w.nb + 15) / 16;
    w.cl = (base + core + (int)window_smem_bytes(w.nb, W) + cl_off) / 4;
    w.cells = 0;
    w.rho_box = 0.0f;
    const int extra = per * GPC;
    w.cfg = extra / 8;
    w.itab = (extra + (int)((sizeof(hans_cfg) + 15) & ~(size_t)15)) / 2;
    // CTA-wide: copy of the configuration and the intra-chain pair table (ordered by b, then j)
    {
        const double *src = reinterpret_cast<const double *>(&c);
        for (int i = threadIdx.x; i < (int)(sizeof(hans_cfg) / 8); i += blockDim.x) smem[w.cfg + i] = src[i];
        uchar2 *tab = reinterpret_cast<uchar2 *>(smem) + w.itab;
        for (int e = threadIdx.x; e < w.npc; e += blockDim.x) {
            int b = w.nex + 1, acc = 0;
            while (acc + (b - w.nex) <= e) { acc += b - w.nex; ++b; }
            tab[e] = make_uchar2((unsigned char)(e - acc), (unsigned char)b);
        }
    }
    __syncthreads();
    return w;
}

template <int T> __device__ __forceinline__ WCtx setup_ctx(const hans_cfg &c, int tail = 0)
{
    return setup_ctx_g<T, cta_threads<T>() / T>(c, 0, tail);
}

// ---------------------------------------------------------------------------------------------
// float32 pre-filter of the pair test.
//
// The decision "r^2 < 1" is always the float64 one (pair_r2, the arithmetic of the CPU checker).
// To reach it cheaply, every bead also has a float32 shadow: its fractional coordinates wrapped
// to [-0.5, 0.5].  For a pair, ds = s_j - s_i is wrapped again and r^2 = ds.G.ds is evaluated
// with the metric tensor G = H H^T in float32 (21 operations instead of 32 float64 ones).  If the
// result is outside the band [1 - eps, 1 + eps] the float64 test is guaranteed to agree and is
// skipped; pairs inside the band (probability ~1e-4) are re-tested in float64.
//
// eps bounds |r2_f32 - r2_f64| for pairs near contact: inputs carry <= 1.5 * 2^-24 per fractional
// component (two stored roundings + one subtraction), which moves r^2 by <= 2 |G ds| . |d(ds)|
// <= 9 Lmax 2^-24 |d|; the nine float32 operations and the rounding of G add <= ~10 * 2^-24 times
// the sum of |terms| <= (sum_i L_i / w_i)^2 |d|^2 (L_i cell vector lengths, w_i perpendicular widths).
// eps = 2^-24 (8 (1 + Lmax^2) + 144 (Lmax / wmin)^2 + 64) * 1.01 uses square-root free upper bounds of
// those terms and carries a further safety factor.  The filter is
// switched off (every pair is re-tested in float64) when a perpendicular width is <= 2.1, where
// float32 and float64 could wrap a pair to different images that both lie within reach of sigma.
// SF layout: g00, g11, g22, 2 g01, 2 g02, 2 g12, lo = 1 - eps, hi = 1 + eps, hw = (half the smallest
// perpendicular width, rounded down): separations below hw are wrapped to their nearest image.
// ---------------------------------------------------------------------------------------------
__device__ __noinline__ void filter_params(int sh, int sf, bool writer)
{
    const double *H = smem + sh, *Hi = smem + sh + 9;
    const double g00 = dot3(H, H), g11 = dot3(H + 3, H + 3), g22 = dot3(H + 6, H + 6);
    const double g01 = dot3(H, H + 3), g02 = dot3(H, H + 6), g12 = dot3(H + 3, H + 6);
    const double b0 = (Hi[0] * Hi[0] + Hi[3] * Hi[3]) + Hi[6] * Hi[6]; // 1 / w_0^2
    const double b1 = (Hi[1] * Hi[1] + Hi[4] * Hi[4]) + Hi[7] * Hi[7];
    const double b2 = (Hi[2] * Hi[2] + Hi[5] * Hi[5]) + Hi[8] * Hi[8];
    const double L2 = fmax(g00, fmax(g11, g22));   // Lmax^2
    const double B2 = fmax(b0, fmax(b1, b2));      // 1 / wmin^2
    // square-root free upper bounds: Lmax <= (1 + Lmax^2)/2, q = sum L_i/w_i <= 3 Lmax/wmin
    const double eps = 5.9604644775390625e-8 * ((8.0 * (1.0 + L2) + 144.0 * (L2 * B2)) + 64.0) * 1.01;
    float lo = (float)(1.0 - eps), hi = (float)(1.0 + eps);
    if (!(B2 * 4.41 < 1.0) || !(eps < 0.05)) { lo = -1.0f; hi = 3.0e38f; } // filter off
    if (writer) {
        float *o = SMF + sf;
        o[0] = (float)g00; o[1] = (float)g11; o[2] = (float)g22;
        o[3] = (float)(2.0 * g01); o[4] = (float)(2.0 * g02); o[5] = (float)(2.0 * g12);
        o[6] = lo; o[7] = hi;
        o[8] = (float)((0.5 / sqrt(B2)) * (1.0 - 1e-6));
    }
}

struct Filt { float g00, g11, g22, g01, g02, g12, lo, hi, hw; };
__device__ __forceinline__ Filt load_filt(int sf)
{
    const float4 a = SM4[sf / 4], b = SM4[sf / 4 + 1];
    Filt f;
    f.g00 = a.x; f.g11 = a.y; f.g22 = a.z; f.g01 = a.w; f.g02 = b.x; f.g12 = b.y; f.lo = b.z; f.hi = b.w;
    f.hw = SMF[sf + 8];
    return f;
}

__device__ __forceinline__ float rint32(float x)
{
    const float magic = 12582912.0f; // 1.5 * 2^23
    return __fsub_rn(__fadd_rn(x, magic), magic);
}

// float32 r^2 of a fractional difference (wrapped here)
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
__device__ __forceinline__ float4 shadow_of(double x, double y, double z, const double *Hi)
{
    double s0, s1, s2;
    vecmat(x, y, z, Hi, s0, s1, s2);
    return make_float4((float)(s0 - hd_rint(s0)), (float)(s1 - hd_rint(s1)), (float)(s2 - hd_rint(s2)), 0.0f);
}

// ---------------------------------------------------------------------------------------------
// single-chain moves: the hot path
// ---------------------------------------------------------------------------------------------
struct Dec { int accepted, reason; double boltz; };
__device__ __forceinline__ Dec mkdec(int a, int r, double b) { Dec d; d.accepted = a; d.reason = r; d.boltz = b; return d; }

// float64 re-test of the pairs of the trial chain that fell into the band (cold)
template <int T>
__device__ __noinline__ bool recheck_chain_exact(WCtx w, int ichain, int b0, bool amb_inter, bool amb_intra)
{
    const int nb = w.nb, c0 = ichain * nb, sh = W_SH(w), cn = W_CN(w), p = W_P(w);
    bool ov = false;
    if (w.lane == 0) atomicAdd(&g_diag[0], 1ull);
    if (amb_inter) {
#pragma unroll 1
        for (int j = w.lane; j < w.N; j += T) {
            if ((unsigned)(j - c0) < (unsigned)nb) continue;
#pragma unroll 1
            for (int b = b0; b < nb; ++b)
                ov |= pair_r2(smem + sh, smem + sh + 9, smem[p + 3 * j] - smem[cn + 3 * b],
                              smem[p + 3 * j + 1] - smem[cn + 3 * b + 1], smem[p + 3 * j + 2] - smem[cn + 3 * b + 2]) < 1.0;
        }
    }
    if (amb_intra) {
        const uchar2 *tab = W_ITAB(w);
#pragma unroll 1
        for (int e = w.lane; e < w.npc; e += T) {
            const uchar2 jb = tab[e];
            if ((int)jb.y >= b0)
                ov |= pair_r2(smem + sh, smem + sh + 9, smem[cn + 3 * jb.y] - smem[cn + 3 * jb.x],
                              smem[cn + 3 * jb.y + 1] - smem[cn + 3 * jb.x + 1], smem[cn + 3 * jb.y + 2] - smem[cn + 3 * jb.x + 2]) < 1.0;
        }
    }
    return ov;
}

// Does the trial chain (CN / C4) overlap anything?  beads [b0, nb) against every bead of every
// other chain, and (dihedral only) the intra-chain pairs (j, b), b >= b0, beyond the exclusion.
template <int T>
__device__ __forceinline__ bool chain_overlaps(const WCtx &w, int ichain, int b0, bool intra, unsigned &pairs)
{
    const Filt f = load_filt(W_SF(w));
    const int nb = w.nb, c0 = ichain * nb, f4 = W_F4(w), c4 = W_C4(w);
    bool ov = false, amb = false;
    for (int j0 = 0; j0 < w.N; j0 += T) {
        const int j = j0 + w.lane;
        if (j < w.N && (unsigned)(j - c0) >= (unsigned)nb) {
            const float4 s = SM4[f4 + j];
            for (int b = b0; b < nb; ++b) {
                const float4 c = SM4[c4 + b];
                const float r2 = r2_f32(f, s.x - c.x, s.y - c.y, s.z - c.z);
                ov |= (r2 < f.lo);
                amb |= !(r2 > f.hi);
            }
            pairs += (unsigned)(nb - b0);
        }
        if (j0 + T < w.N && gany<T>(ov)) return true; // early exit between passes
    }
    bool amb_intra = false;
    if (intra) {
        const uchar2 *tab = W_ITAB(w);
        for (int e = w.lane; e < w.npc; e += T) {
            const uchar2 jb = tab[e];
            if ((int)jb.y >= b0) {
                const float4 p = SM4[c4 + jb.x], q = SM4[c4 + jb.y];
                const float r2 = r2_f32(f, q.x - p.x, q.y - p.y, q.z - p.z);
                ov |= (r2 < f.lo);
                amb_intra |= !(r2 > f.hi);
                pairs += 1;
            }
        }
    }
    if (gany<T>(ov)) return true;
    // nothing overlaps for certain; any pair inside the band decides in float64
    const bool any_inter = gany<T>(amb), any_intra = intra && gany<T>(amb_intra);
    if (any_inter || any_intra)
        return gany<T>(recheck_chain_exact<T>(w, ichain, b0, amb, amb_intra));
    return false;
}

// Trial chain of a translate / rotate / dihedral proposal, built from the CURRENT chain: beads
// first, first+stride, ... are written to smem[cn + 3b ..] (float64) and SM4[c4 + b] (shadows).
// Returns false for a malformed dihedral proposal.  One copy of this arithmetic serves both the
// cooperative path (first = lane, stride = T) and the speculative windows (hans_lean.cuh).
__device__ __forceinline__ bool make_trial(const WCtx &w, const PropRec *p, int first, int stride, int cn, int c4,
                                           int &b0, bool &intra)
{
    const int nb = w.nb, type = p->type;
    const int P = W_P(w) + 3 * p->ichain * nb, sh = W_SH(w);
    b0 = 0;
    intra = false;
    if (type == HANS_TRANS) {
        for (int b = first; b < nb; b += stride) {
            const double x = smem[P + 3 * b] + p->p[0], y = smem[P + 3 * b + 1] + p->p[1], z = smem[P + 3 * b + 2] + p->p[2];
            smem[cn + 3 * b] = x; smem[cn + 3 * b + 1] = y; smem[cn + 3 * b + 2] = z;
            SM4[c4 + b] = shadow_of(x, y, z, smem + sh + 9);
        }
    } else if (type == HANS_ROT) {
        double com[3] = {0.0, 0.0, 0.0};
        for (int b = 0; b < nb; ++b) { com[0] += smem[P + 3 * b]; com[1] += smem[P + 3 * b + 1]; com[2] += smem[P + 3 * b + 2]; }
        const double dn = (double)nb;
        com[0] /= dn; com[1] /= dn; com[2] /= dn;
        for (int b = first; b < nb; b += stride) {
            const double v[3] = {smem[P + 3 * b] - com[0], smem[P + 3 * b + 1] - com[1], smem[P + 3 * b + 2] - com[2]};
            double o[3];
            quat_rot(p->p, v, o);
            const double x = o[0] + com[0], y = o[1] + com[1], z = o[2] + com[2];
            smem[cn + 3 * b] = x; smem[cn + 3 * b + 1] = y; smem[cn + 3 * b + 2] = z;
            SM4[c4 + b] = shadow_of(x, y, z, smem + sh + 9);
        }
    } else {
        const int ia = p->idx0;
        if (nb < 4 || ia < 3 || ia > nb - 1) return false;
        const bool flip = p->idx1 != 0;
        const int iv = P + 3 * (flip ? nb - 1 - (ia - 1) : ia - 1);
        const int iw = P + 3 * (flip ? nb - 1 - (ia - 2) : ia - 2);
        const double piv[3] = {smem[iv], smem[iv + 1], smem[iv + 2]};
        double ax[3] = {piv[0] - smem[iw], piv[1] - smem[iw + 1], piv[2] - smem[iw + 2]};
        const double n = sqrt(dot3(ax, ax));
        ax[0] /= n; ax[1] /= n; ax[2] /= n;
        double s, co;
        hd_sincos(0.5 * p->p[0], &s, &co);
        const double q[4] = {co, s * ax[0], s * ax[1], s * ax[2]};
        for (int b = first; b < nb; b += stride) {
            const int src = P + 3 * (flip ? nb - 1 - b : b);
            double r[3] = {smem[src], smem[src + 1], smem[src + 2]};
            if (b >= ia) {
                const double v[3] = {r[0] - piv[0], r[1] - piv[1], r[2] - piv[2]};
                double o[3];
                quat_rot(q, v, o);
                r[0] = o[0] + piv[0]; r[1] = o[1] + piv[1]; r[2] = o[2] + piv[2];
            }
            smem[cn + 3 * b] = r[0]; smem[cn + 3 * b + 1] = r[1]; smem[cn + 3 * b + 2] = r[2];
            SM4[c4 + b] = shadow_of(r[0], r[1], r[2], smem + sh + 9);
        }
        b0 = ia;
        intra = true;
    }
    return true;
}

// ---------------------------------------------------------------------------------------------
// Chain-level pruning (chains of more than one bead).
//
// Every chain c carries a bounding sphere: centre = its middle bead (whose float32 shadow is already
// in F4 / G4 / C4), radius RAD[c] = the largest float64 distance of a bead from it, rounded up.  Two
// chains can only have a bead pair closer than sigma = 1 if their centres are closer than
// D = R_a + R_b + 1, so a pair of chains whose float32 centre separation exceeds D by a margin
// (0.001 in D and 1e-5 relative in D^2, orders of magnitude above the float32 error of the shadows
// and of the metric) is skipped WITHOUT testing its beads; only the chains that survive go through
// the bead-level float32 filter / float64 settle above.  The test is valid while D is below half the
// smallest perpendicular width of the cell (`hw`): then the wrapped fractional separation of the
// centres is their nearest image, and so is that of every bead pair within sigma.  For smaller
// cells nothing is skipped.  Skipped pairs are certain non-overlaps, so decisions -- and with them
// whole trajectories -- are unchanged; only the number of executed pair tests drops (hexamers in
// the dilute regime: 540 -> ~60 per chain move).  This is the role hs_alkane's link cells play
// (bypassed by the reference, MCNS.py:603), at chain granularity.
// ---------------------------------------------------------------------------------------------
#ifndef HANS_BEAD_PRUNE
#define HANS_BEAD_PRUNE 1
#endif
__device__ __forceinline__ float chain_radius(int pos, int nb)
{
    const int m = nb >> 1;
    const double mx = smem[pos + 3 * m], my = smem[pos + 3 * m + 1], mz = smem[pos + 3 * m + 2];
    double r2 = 0.0;
#pragma unroll 1
    for (int b = 0; b < nb; ++b) {
        const double dx = smem[pos + 3 * b] - mx, dy = smem[pos + 3 * b + 1] - my, dz = smem[pos + 3 * b + 2] - mz;
        const double d2 = (dx * dx + dy * dy) + dz * dz;
        r2 = d2 > r2 ? d2 : r2;
    }
    return __fsqrt_ru(__double2float_ru(r2));
}

__device__ __forceinline__ bool spheres_near(const Filt &f, const float4 a, const float4 b, float D)
{
    const float r2c = r2_f32(f, a.x - b.x, a.y - b.y, a.z - b.z);
    return !(D < f.hw) || !(r2c > D * D * 1.00001f);
}

// Compaction of the items a wave of T threads flags `near` into LST; returns their number.
// Collective over the group; the group must synchronise (gsync / gany) before the next call.
template <int T>
__device__ __forceinline__ int near_list(const WCtx &w, bool near, int item, bool counter_is_zero = false)
{
    int *lst = reinterpret_cast<int *>(smem) + w.lst;
    if (T == 32) {
        const unsigned mask = __ballot_sync(FULL, near);
        if (near) lst[__popc(mask & ((1u << w.lane) - 1u))] = item;
        __syncwarp();
        return __popc(mask);
    } else {
        int *cnt = lst - 4;
        if (!counter_is_zero) { // (else the caller zeroed it before its last barrier: one barrier less)
            if (w.lane == 0) *cnt = 0;
            __syncthreads();
        }
        if (near) lst[atomicAdd(cnt, 1)] = item;
        __syncthreads();
        return *cnt;
    }
}

// One CTA per walker (T > 32): near-list counter and vote word are zeroed by the caller before the barrier
// that publishes the trial chain, so a chain move costs four barriers (trial, near list, vote, commit)
template <int T> __device__ __forceinline__ void coop_reset(const WCtx &w)
{
    if (T > 32 && w.lane == 0) {
        int *z = reinterpret_cast<int *>(smem) + w.lst - 4;
        z[0] = 0; z[1] = 0;
    }
}

// ---------------------------------------------------------------------------------------------
// Per-walker cell list (layout at cell_smem_bytes).
//
// Correctness: the decision is the single-image float64 r^2 < 1 (pair_r2).  For such a pair the wrapped fractional
// difference ds = d.H^-1 obeys |ds_k| <= |d| / w_k < 1 / w_k, and a bin is at least 1.001 / w_k wide, so the two beads sit
// in the same or in adjacent bins (periodically) along every axis -- also when the wrapped image is not the nearest one.
// Bins are computed from the float32 shadows (wrapped fractional coordinates, error 2^-24): the 0.1 % margin of the
// bin width covers that for every cell narrower than ~10^4 sigma.  So the beads found in the <= 27 bins around a
// bead are a superset of its overlap partners: decisions are those of the all-pairs test, bit for bit.  Axes with fewer
// than three bins enumerate each bin once.
// ---------------------------------------------------------------------------------------------
struct Grid { int n0, n1, n2; };
__device__ __forceinline__ Grid load_grid(int dims) { Grid g; g.n0 = SMI[dims]; g.n1 = SMI[dims + 1]; g.n2 = SMI[dims + 2]; return g; }
__device__ __forceinline__ int cl_coord(float s, int n)
{
    int b = (int)((s + 0.5f) * (float)n);
    b = b < 0 ? 0 : b;
    return b >= n ? n - 1 : b;
}
__device__ __forceinline__ int cl_bin(const Grid &g, const float4 s)
{
    return (cl_coord(s.x, g.n0) * g.n1 + cl_coord(s.y, g.n1)) * g.n2 + cl_coord(s.z, g.n2);
}
// neighbour o (0 .. min(n,3)-1) of coordinate b along an axis of n bins
__device__ __forceinline__ int cl_nbr(int b, int o, int n)
{
    int x = b + o - (n >= 3 ? 1 : 0);
    x = x < 0 ? x + n : x;
    return x >= n ? x - n : x;
}

// grid of the cell whose inverse is at smem[hi]; written to dims by one thread
__device__ __noinline__ void cl_dims(int hi, int dims)
{
    const double *Hi = smem + hi;
    const double b0 = (Hi[0] * Hi[0] + Hi[3] * Hi[3]) + Hi[6] * Hi[6]; // 1 / w_0^2
    const double b1 = (Hi[1] * Hi[1] + Hi[4] * Hi[4]) + Hi[7] * Hi[7];
    const double b2 = (Hi[2] * Hi[2] + Hi[5] * Hi[5]) + Hi[8] * Hi[8];
    int n[3];
    const double bb[3] = {b0, b1, b2};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double wk = 1.0 / sqrt(bb[k]);
        double q = floor(wk / 1.001);
        q = q < 1.0 ? 1.0 : (q > (double)CL_M ? (double)CL_M : q);
        n[k] = (int)q;
    }
    SMI[dims] = n[0]; SMI[dims + 1] = n[1]; SMI[dims + 2] = n[2]; SMI[dims + 3] = n[0] * n[1] * n[2];
}

// (re)build a list from the shadows at q4 under the inverse cell at smem[hi]; collective, ends with a group barrier
template <int T>
__device__ __noinline__ void cl_build(WCtx w, int hi, int q4, int dims, int head, int next)
{
    if (w.lane == 0) cl_dims(hi, dims);
    gsync<T>();
    const Grid g = load_grid(dims);
    const int nbins = g.n0 * g.n1 * g.n2;
#pragma unroll 1
    for (int b = w.lane; b < nbins; b += T) SMI[head + b] = -1;
    gsync<T>();
#pragma unroll 1
    for (int i = w.lane; i < w.N; i += T) SMI[next + i] = atomicExch(SMI + head + cl_bin(g, SM4[q4 + i]), i);
    gsync<T>();
}

// the chain's beads leave the bins of their old shadows (F4) and enter those of the trial shadows (C4) -- ALL of them:
// a dihedral move with a flip relabels the beads it does not move.  Called by ONE WARP (all 32 lanes): the bins are
// computed in parallel, then lane 0 relinks the beads whose bin changed one after the other (concurrent unlinking is
// not safe; in the dense regime most beads stay in their bin).  Call BEFORE F4 is overwritten.
__device__ __noinline__ void cl_move_chain(WCtx w, int ichain)
{
    const Grid g = load_grid(CL_DIMS(w));
    const int head = CL_HEAD(w), next = CL_NEXT(w), c0 = ichain * w.nb, ln = threadIdx.x & 31;
#pragma unroll 1
    for (int p0 = 0; p0 < w.nb; p0 += 32) {
        const int b = p0 + ln;
        int bo = 0, bn = 0;
        if (b < w.nb) { bo = cl_bin(g, SM4[W_F4(w) + c0 + b]); bn = cl_bin(g, SM4[W_C4(w) + b]); }
        unsigned mask = __ballot_sync(FULL, bo != bn);
        while (mask) {
            const int l = __ffs(mask) - 1;
            mask &= mask - 1u;
            const int fo = __shfl_sync(FULL, bo, l), tn = __shfl_sync(FULL, bn, l);
            if (ln == 0) {
                const int i = c0 + p0 + l;
                int j = SMI[head + fo], prev = -1;
                for (int guard = 0; j != i && j >= 0 && guard < w.N; ++guard) { prev = j; j = SMI[next + j]; }
                if (j == i) {
                    if (prev < 0) SMI[head + fo] = SMI[next + i]; else SMI[next + prev] = SMI[next + i];
                }
                SMI[next + i] = SMI[head + tn];
                SMI[head + tn] = i;
            }
        }
    }
    __syncwarp();
}

// Does the trial chain (CN / C4, bounding radius Rt) overlap anything?  NB > 0: beads per chain known
// at compile time, trial shadows in registers.
template <int T, int NB>
__device__ __forceinline__ bool chain_overlaps_pruned(const WCtx &w, const Filt &f, int ichain, int b0, bool intra,
                                                      float Rt, unsigned &pairs)
{
    const int nb = NB ? NB : w.nb, nc = w.nc, m = nb >> 1, f4 = W_F4(w), c4 = W_C4(w);
    const int *lst = reinterpret_cast<const int *>(smem) + w.lst;
    float4 tr[NB ? NB : 1];
    if (NB) {
#pragma unroll
        for (int b = 0; b < NB; ++b) tr[b] = SM4[c4 + b];
    }
    const float4 ct = SM4[c4 + m];
    bool ov = false, amb = false;
    for (int cb = 0; cb < nc; cb += T) {
        const int c = cb + w.lane;
        bool near = false;
        if (c < nc && c != ichain) near = spheres_near(f, SM4[f4 + c * nb + m], ct, Rt + SMF[w.rad + c] + 1.001f);
        const int items = near_list<T>(w, near, c, cb == 0) * nb;
#if HANS_BEAD_PRUNE
        if (T == 32 && w.N >= 32 && items <= 32) {
            // few near chains (the usual case outside the dense regime; measured: with more than one pass of beads the
            // unrolled loop below, six independent tests in flight per lane, is faster although it executes more tests).
            // Bead level in two steps (one warp per walker): which beads of the near chains lie inside the trial
            // chain's bounding sphere (+ sigma) -- one test per bead, the survivors' shadows go to a 32-entry buffer
            // (the head of G4: the trial shadows of a box move, free during a chain move); then the (bead, trial bead)
            // pairs of the survivors spread over the lanes.  Beads outside the sphere are further than sigma from every
            // trial bead, so the decision is unchanged.
            float4 *buf = SM4 + W_G4(w);
#pragma unroll 1
            for (int e0 = 0; e0 < items; e0 += 32) {
                const int e = e0 + w.lane;
                bool in = false;
                float4 s = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                if (e < items) {
                    const int ci = NB ? e / (NB ? NB : 1) : div_nb(w, e);
                    s = SM4[f4 + lst[ci] * nb + (e - ci * nb)];
                    in = spheres_near(f, s, ct, Rt + 1.001f);
                    pairs += 1u;
                }
                const unsigned mask = __ballot_sync(FULL, in);
                if (in) buf[__popc(mask & ((1u << w.lane) - 1u))] = s;
                __syncwarp();
                const int npair = __popc(mask) * nb;
#pragma unroll 1
                for (int t = w.lane; t < npair; t += 32) {
                    const int bi = NB ? t / (NB ? NB : 1) : div_nb(w, t);
                    const int b = t - bi * nb;
                    const float4 sb = buf[bi], tb = SM4[c4 + b];
                    const float r2 = r2_f32(f, sb.x - tb.x, sb.y - tb.y, sb.z - tb.z);
                    const bool moved = b >= b0;
                    ov |= moved && (r2 < f.lo);
                    amb |= moved && !(r2 > f.hi);
                    pairs += moved ? 1u : 0u;
                }
                __syncwarp(); // buf is rewritten by the next pass
            }
        } else
#endif
        for (int e = w.lane; e < items; e += T) {
            const int ci = NB ? e / (NB ? NB : 1) : div_nb(w, e);
            const float4 s = SM4[f4 + lst[ci] * nb + (e - ci * nb)];
            if (NB) {
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    const float r2 = r2_f32(f, s.x - tr[b].x, s.y - tr[b].y, s.z - tr[b].z);
                    const bool moved = b >= b0;
                    ov |= moved && (r2 < f.lo);
                    amb |= moved && !(r2 > f.hi);
                }
            } else {
                for (int b = b0; b < nb; ++b) {
                    const float4 t = SM4[c4 + b];
                    const float r2 = r2_f32(f, s.x - t.x, s.y - t.y, s.z - t.z);
                    ov |= (r2 < f.lo);
                    amb |= !(r2 > f.hi);
                }
            }
            pairs += (unsigned)(nb - b0);
        }
        pairs += (c < nc) ? 1u : 0u;
        if (cb + T < nc) gsync<T>(); // LST is rewritten by the next wave
    }
    bool amb_intra = false;
    if (intra) {
        const uchar2 *tab = W_ITAB(w);
        for (int e = w.lane; e < w.npc; e += T) {
            const uchar2 jb = tab[e];
            if ((int)jb.y >= b0) {
                const float4 p = SM4[c4 + jb.x], q = SM4[c4 + jb.y];
                const float r2 = r2_f32(f, q.x - p.x, q.y - p.y, q.z - p.z);
                ov |= (r2 < f.lo);
                amb_intra |= !(r2 > f.hi);
                pairs += 1;
            }
        }
    }
    if (T == 32) {
        const unsigned code = __reduce_or_sync(FULL, (ov ? 4u : 0u) | (amb ? 1u : 0u) | (amb_intra ? 2u : 0u));
        if (code & 4u) return true;
        if (code) return __any_sync(FULL, recheck_chain_exact<T>(w, ichain, b0, (code & 1u) != 0u, (code & 2u) != 0u));
        return false;
    }
    {   // T > 32: one barrier for the three flags (warp REDUX, one atomicOr per warp into the zeroed vote word)
        int *vote = reinterpret_cast<int *>(smem) + w.lst - 3;
        const unsigned wcode = __reduce_or_sync(FULL, (ov ? 4u : 0u) | (amb ? 1u : 0u) | (amb_intra ? 2u : 0u));
        if ((threadIdx.x & 31) == 0 && wcode) atomicOr(vote, (int)wcode);
        __syncthreads();
        const unsigned code = (unsigned)*vote;
        if (code & 4u) return true;
        if (code) return gany<T>(recheck_chain_exact<T>(w, ichain, b0, (code & 1u) != 0u, (code & 2u) != 0u));
        return false;
    }
}

// Does the trial chain (CN / C4) overlap anything?  Cell-list version: one task per (moved trial bead, neighbour bin),
// spread over the group's threads; every bead found there that belongs to another chain goes through the float32
// filter.  Same votes and the same float64 settle as chain_overlaps_pruned.  G3: at least three bins along every axis
// (27 distinct neighbours; the task index is decoded with compile-time divisions).
template <int T, bool G3>
__device__ __forceinline__ void chain_cells_scan(const WCtx &w, const Filt &f, const Grid &g, int ichain, int b0, bool &ov, bool &amb,
                                                 unsigned &pairs)
{
    const int nb = w.nb, c0 = ichain * nb, f4 = W_F4(w), c4 = W_C4(w), head = CL_HEAD(w), next = CL_NEXT(w);
    const int k0 = G3 ? 3 : (g.n0 < 3 ? g.n0 : 3), k1 = G3 ? 3 : (g.n1 < 3 ? g.n1 : 3), k2 = G3 ? 3 : (g.n2 < 3 ? g.n2 : 3);
    const int k12 = k1 * k2, K = k0 * k12, ntask = (nb - b0) * K;
#pragma unroll 1
    for (int t = w.lane; t < ntask; t += T) {
        const int bi = t / K, o = t - bi * K;
        const int o0 = o / k12, o1 = (o - o0 * k12) / k2, o2 = o - o0 * k12 - o1 * k2;
        const float4 tb = SM4[c4 + b0 + bi];
        const int bin = (cl_nbr(cl_coord(tb.x, g.n0), o0, g.n0) * g.n1 + cl_nbr(cl_coord(tb.y, g.n1), o1, g.n1)) * g.n2 +
                        cl_nbr(cl_coord(tb.z, g.n2), o2, g.n2);
#pragma unroll 1
        for (int j = SMI[head + bin], guard = 0; j >= 0 && guard < w.N; j = SMI[next + j], ++guard) {
            if ((unsigned)(j - c0) < (unsigned)nb) continue;
            const float4 sj = SM4[f4 + j];
            const float r2 = r2_f32(f, sj.x - tb.x, sj.y - tb.y, sj.z - tb.z);
            ov |= (r2 < f.lo);
            amb |= !(r2 > f.hi);
            pairs += 1u;
        }
    }
}

template <int T>
__device__ __forceinline__ bool chain_overlaps_cells(const WCtx &w, const Filt &f, int ichain, int b0, bool intra, unsigned &pairs)
{
    const Grid g = load_grid(CL_DIMS(w));
    const int c4 = W_C4(w);
    bool ov = false, amb = false;
    if (g.n0 >= 3 && g.n1 >= 3 && g.n2 >= 3) chain_cells_scan<T, true>(w, f, g, ichain, b0, ov, amb, pairs);
    else chain_cells_scan<T, false>(w, f, g, ichain, b0, ov, amb, pairs);
    bool amb_intra = false;
    if (intra) {
        const uchar2 *tab = W_ITAB(w);
        for (int e = w.lane; e < w.npc; e += T) {
            const uchar2 jb = tab[e];
            if ((int)jb.y >= b0) {
                const float4 p = SM4[c4 + jb.x], q = SM4[c4 + jb.y];
                const float r2 = r2_f32(f, q.x - p.x, q.y - p.y, q.z - p.z);
                ov |= (r2 < f.lo);
                amb_intra |= !(r2 > f.hi);
                pairs += 1;
            }
        }
    }
    if (T == 32) {
        const unsigned code = __reduce_or_sync(FULL, (ov ? 4u : 0u) | (amb ? 1u : 0u) | (amb_intra ? 2u : 0u));
        if (code & 4u) return true;
        if (code) return __any_sync(FULL, recheck_chain_exact<T>(w, ichain, b0, (code & 1u) != 0u, (code & 2u) != 0u));
        return false;
    }
    {   // T > 32: one barrier for the three flags (the caller zeroed the vote word, coop_reset)
        int *vote = reinterpret_cast<int *>(smem) + w.lst - 3;
        const unsigned wcode = __reduce_or_sync(FULL, (ov ? 4u : 0u) | (amb ? 1u : 0u) | (amb_intra ? 2u : 0u));
        if ((threadIdx.x & 31) == 0 && wcode) atomicOr(vote, (int)wcode);
        __syncthreads();
        const unsigned code = (unsigned)*vote;
        if (code & 4u) return true;
        if (code) return gany<T>(recheck_chain_exact<T>(w, ichain, b0, (code & 1u) != 0u, (code & 2u) != 0u));
        return false;
    }
}

// translate / rotate / dihedral: alk.alkane_translate_chain, alkane_rotate_chain,
// alkane_bond_rotate (MCNS.py:323,328,333) + accept/restore (MCNS.py:371-380)
template <int T>
__device__ __forceinline__ Dec move_chain(const WCtx &w, const PropRec *p, int mode, unsigned &pairs)
{
    const int nb = w.nb, ichain = p->ichain;
    const int c0 = ichain * nb, P = W_P(w) + 3 * c0, cn = W_CN(w), c4 = W_C4(w);
    int b0 = 0;
    bool intra = false;
    if (!make_trial(w, p, w.lane, T, cn, c4, b0, intra)) return mkdec(0, HANS_REJ_INVALID, 0.0);
    coop_reset<T>(w);
    gsync<T>();
    bool ov;
    float Rt = 0.0f;
    if (nb > 1) Rt = intra ? chain_radius(cn, nb) : SMF[w.rad + ichain]; // translations and rotations are rigid
    if (w.cells & 1) {
        ov = chain_overlaps_cells<T>(w, load_filt(W_SF(w)), ichain, b0, intra, pairs);
    } else if (nb > 1) {
        ov = chain_overlaps_pruned<T, 0>(w, load_filt(W_SF(w)), ichain, b0, intra, Rt, pairs);
    } else {
        ov = chain_overlaps<T>(w, ichain, b0, intra, pairs);
    }
    const double boltz = ov ? 0.0 : 1.0;
    if ((mode == HANS_MODE_PROPOSE) || (p->u_acc < boltz)) { // MCNS.py:372
        if (w.cells & 1) { // (every thread is past its reads of the list: the vote above was a group barrier)
            if (w.lane < 32) cl_move_chain(w, ichain);
            gsync<T>();
        }
        const int f4 = W_F4(w) + c0;
        for (int b = w.lane; b < nb; b += T) {
            smem[P + 3 * b] = smem[cn + 3 * b]; smem[P + 3 * b + 1] = smem[cn + 3 * b + 1]; smem[P + 3 * b + 2] = smem[cn + 3 * b + 2];
            SM4[f4 + b] = SM4[c4 + b];
        }
        if (intra && w.lane == 0) SMF[w.rad + ichain] = Rt;
    }
    gsync<T>();
    return mkdec(ov ? 0 : 1, ov ? HANS_REJ_OVERLAP : HANS_ACCEPT, boltz);
}

// ---------------------------------------------------------------------------------------------
// box moves (cold: out of line, context by value)
// ---------------------------------------------------------------------------------------------
// all pairs of the beads at `pos` in float64 (the reference arithmetic), bead-level round robin:
// lane owns bead i and walks the offsets d = 1..N/2 with j = (i + d) mod N, so every unordered pair
// is visited once (for even N the antipodal offset is taken by the lower half only).  Same-chain
// pairs are skipped unless with_intra and further apart than nexclude.
template <int T>
__device__ __forceinline__ bool all_overlap64(const WCtx &w, int pos, const double *H, const double *Hi, bool with_intra)
{
    const int nb = w.nb, N = w.N, half = N >> 1;
    const bool even = (N & 1) == 0;
    const int RB = T < N ? T : N;
    const int DG = T < N ? 1 : T / N;
    const int li = w.lane % RB, dg = w.lane / RB;
    const bool lane_on = dg < DG;
    bool ov = false;
    constexpr int U = 1;
    for (int i0 = 0; i0 < N; i0 += RB) {
        const int i = i0 + li;
        const bool row_on = lane_on && i < N;
        const int ii = row_on ? i : 0;
        const double xi = smem[pos + 3 * ii], yi = smem[pos + 3 * ii + 1], zi = smem[pos + 3 * ii + 2];
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
                    if (!same || (with_intra && sep > w.nex))
                        ov |= (pair_r2(H, Hi, smem[pos + 3 * j] - xi, smem[pos + 3 * j + 1] - yi, smem[pos + 3 * j + 2] - zi) < 1.0);
                }
            }
            if (gany<T>(ov)) return true;
        }
    }
    return false;
}

// float64 test of the TRIAL beads under the trial cell (cold: band hits, single-chain systems)
template <int T>
__device__ __noinline__ bool trial_overlap_exact(WCtx w, bool with_intra)
{
    double H[9], Hi[9];
    if (w.lane == 0) atomicAdd(&g_diag[1], 1ull);
#pragma unroll
    for (int k = 0; k < 9; ++k) { H[k] = smem[W_SH(w) + 18 + k]; Hi[k] = smem[W_SH(w) + 27 + k]; }
    return all_overlap64<T>(w, W_TP(w), H, Hi, with_intra);
}

// All pairs of the TRIAL beads (TP / G4 under the trial cell) through a cell list built for the trial cell: one task
// per (bead i, neighbour bin).  G3 (at least three bins along every axis): the bead's own bin (pairs j > i) and the 13
// "forward" neighbours, so every pair of beads meets once; otherwise every distinct neighbour bin with j > i.  Same-chain
// pairs only if with_intra and further apart than nexclude.  A vote every CL_VOTE tasks per thread ends a rejected move
// early.  Pairs inside the filter's band: the first one per thread is settled in float64, more than one -> the full pass.
#define CL_VOTE 8
template <int T, bool G3>
__device__ __forceinline__ bool trial_cells_scan(const WCtx &w, const Filt &f, const Grid &g, bool with_intra, bool &ov, int &namb,
                                                 int &ai, int &aj, unsigned &pairs)
{
    const int nb = w.nb, N = w.N, g4 = W_G4(w), head = CL_HEAD2(w), next = CL_NEXT2(w);
    const int k0 = G3 ? 3 : (g.n0 < 3 ? g.n0 : 3), k1 = G3 ? 3 : (g.n1 < 3 ? g.n1 : 3), k2 = G3 ? 3 : (g.n2 < 3 ? g.n2 : 3);
    const int k12 = k1 * k2, K = G3 ? 14 : k0 * k12, ntask = N * K;
    const int nex_eff = with_intra ? w.nex : 0x3fffffff;
#pragma unroll 1
    for (int t0 = 0; t0 < ntask; t0 += T * CL_VOTE) {
#pragma unroll 1
        for (int u = 0; u < CL_VOTE; ++u) {
            const int t = t0 + u * T + w.lane;
            if (t < ntask) {
                const int i = t / K, oo = t - i * K;
                const int o = G3 ? oo + 13 : oo; // G3: linear offsets 13 (own bin) .. 26 of the 27
                const int o0 = o / k12, o1 = (o - o0 * k12) / k2, o2 = o - o0 * k12 - o1 * k2;
                const float4 si = SM4[g4 + i];
                const int bin = (cl_nbr(cl_coord(si.x, g.n0), o0, g.n0) * g.n1 + cl_nbr(cl_coord(si.y, g.n1), o1, g.n1)) * g.n2 +
                                cl_nbr(cl_coord(si.z, g.n2), o2, g.n2);
                const int ci0 = div_nb(w, i) * nb;
                const bool ordered = !G3 || oo == 0; // the same bin on both sides: take each pair once
#pragma unroll 1
                for (int j = SMI[head + bin], guard = 0; j >= 0 && guard < N; j = SMI[next + j], ++guard) {
                    if (ordered && j <= i) continue;
                    if ((unsigned)(j - ci0) < (unsigned)nb && (j > i ? j - i : i - j) <= nex_eff) continue; // same chain, excluded
                    const float4 sj = SM4[g4 + j];
                    const float r2 = r2_f32(f, sj.x - si.x, sj.y - si.y, sj.z - si.z);
                    ov |= (r2 < f.lo);
                    if (!(r2 < f.lo) && !(r2 > f.hi)) {
                        if (namb == 0) { ai = i; aj = j; }
                        ++namb;
                    }
                    pairs += 1u;
                }
            }
        }
        if (gany<T>(ov)) return true;
    }
    return false;
}

// The same scan for grids of at least three bins per axis, sixteen lanes per bead: lane s of a group owns ONE of the 14
// forward offsets (s = 0: the bead's own bin) for the whole scan, so no task index is decoded inside the loop.
template <int T>
__device__ __forceinline__ bool trial_cells_scan16(const WCtx &w, const Filt &f, const Grid &g, bool with_intra, bool &ov, int &namb,
                                                   int &ai, int &aj, unsigned &pairs)
{
    const int nb = w.nb, N = w.N, g4 = W_G4(w), head = CL_HEAD2(w), next = CL_NEXT2(w);
    const int sub = w.lane & 15, grp = w.lane >> 4;
    constexpr int BPP = T / 16; // beads per pass of the group
    const int o = sub + 13;     // linear offsets 13 (own bin) .. 26 of the 27
    const int d0 = o / 9 - 1, d1 = (o - (o / 9) * 9) / 3 - 1, d2 = o % 3 - 1;
    const bool act = sub < 14, ordered = sub == 0;
    const int nex_eff = with_intra ? w.nex : 0x3fffffff;
#pragma unroll 1
    for (int i0 = 0; i0 < N; i0 += BPP * CL_VOTE) {
#pragma unroll 1
        for (int u = 0; u < CL_VOTE; ++u) {
            const int i = i0 + u * BPP + grp;
            if (act && i < N) {
                const float4 si = SM4[g4 + i];
                int bx = cl_coord(si.x, g.n0) + d0, by = cl_coord(si.y, g.n1) + d1, bz = cl_coord(si.z, g.n2) + d2;
                bx = bx < 0 ? bx + g.n0 : (bx >= g.n0 ? bx - g.n0 : bx);
                by = by < 0 ? by + g.n1 : (by >= g.n1 ? by - g.n1 : by);
                bz = bz < 0 ? bz + g.n2 : (bz >= g.n2 ? bz - g.n2 : bz);
                int j = SMI[head + (bx * g.n1 + by) * g.n2 + bz];
                if (j >= 0) {
                    const int ci0 = div_nb(w, i) * nb;
#pragma unroll 1
                    for (int guard = 0; j >= 0 && guard < N; j = SMI[next + j], ++guard) {
                        if (ordered && j <= i) continue;
                        if ((unsigned)(j - ci0) < (unsigned)nb && (j > i ? j - i : i - j) <= nex_eff) continue; // same chain, excluded
                        const float4 sj = SM4[g4 + j];
                        const float r2 = r2_f32(f, sj.x - si.x, sj.y - si.y, sj.z - si.z);
                        ov |= (r2 < f.lo);
                        if (!(r2 < f.lo) && !(r2 > f.hi)) {
                            if (namb == 0) { ai = i; aj = j; }
                            ++namb;
                        }
                        pairs += 1u;
                    }
                }
            }
        }
        if (gany<T>(ov)) return true;
    }
    return false;
}

template <int T>
__device__ __forceinline__ bool trial_overlap_cells(const WCtx &w, bool with_intra, unsigned &pairs)
{
    cl_build<T>(w, W_SH(w) + 27, W_G4(w), CL_DIMS(w) + 4, CL_HEAD2(w), CL_NEXT2(w));
    const Filt f = load_filt(W_SF(w) + SF_TRIAL);
    const Grid g = load_grid(CL_DIMS(w) + 4);
    const int tp = W_TP(w), sh = W_SH(w) + 18;
    bool ov = false;
    int namb = 0, ai = 0, aj = 0;
    const bool hit = (g.n0 >= 3 && g.n1 >= 3 && g.n2 >= 3) ? trial_cells_scan16<T>(w, f, g, with_intra, ov, namb, ai, aj, pairs)
                                                           : trial_cells_scan<T, false>(w, f, g, with_intra, ov, namb, ai, aj, pairs);
    if (hit) return true;
    if (gany<T>(namb > 0)) {
        if (namb > 0)
            ov |= pair_r2(smem + sh, smem + sh + 9, smem[tp + 3 * aj] - smem[tp + 3 * ai], smem[tp + 3 * aj + 1] - smem[tp + 3 * ai + 1],
                          smem[tp + 3 * aj + 2] - smem[tp + 3 * ai + 2]) < 1.0;
        if (gany<T>(ov)) return true;
        if (gany<T>(namb > 1)) return trial_overlap_exact<T>(w, with_intra);
    }
    return false;
}

// the trial list becomes the current one (accepted box move)
template <int T> __device__ __forceinline__ void cl_commit_trial(const WCtx &w)
{
    const int n = SMI[CL_DIMS(w) + 7];
#pragma unroll 1
    for (int b = w.lane; b < n; b += T) SMI[CL_HEAD(w) + b] = SMI[CL_HEAD2(w) + b];
#pragma unroll 1
    for (int i = w.lane; i < w.N; i += T) SMI[CL_NEXT(w) + i] = SMI[CL_NEXT2(w) + i];
    if (w.lane < 4) SMI[CL_DIMS(w) + w.lane] = SMI[CL_DIMS(w) + 4 + w.lane];
    gsync<T>();
}

// Box moves of chain systems: all pairs of CHAINS through their bounding spheres (round robin over
// the chains: item p = (row ci, offset d), cj = ci + d mod nc), bead level only for the pairs of
// chains that survive.  Every WARP works through its own waves of 32 chain pairs (ballot compaction into
// its 32 entries of LST, bead tests of the survivors) without a CTA-wide barrier; a warp that finds a
// certain overlap raises a flag the others poll, so a rejected move still ends early.  One vote at the
// end.  Pairs of beads inside the filter's band are settled in float64 at the end (first one per thread;
// more than one -> the full float64 pass).  The caller zeroes the flag before its last barrier (T > 32).
template <int T>
__device__ __forceinline__ bool trial_overlap_pruned(const WCtx &w, bool with_intra, unsigned &pairs)
{
    const Filt f = load_filt(W_SF(w) + SF_TRIAL);
    const int nb = w.nb, nc = w.nc, m = nb >> 1, g4 = W_G4(w), tp = W_TP(w), sh = W_SH(w) + 18;
    const int half = nc >> 1, nitems = nc * half;
    const bool even = (nc & 1) == 0;
    const int ln = w.lane & 31, w0 = w.lane & ~31;
    int *lst = reinterpret_cast<int *>(smem) + w.lst + w0;
    volatile int *flag = reinterpret_cast<volatile int *>(smem) + w.lst - 3;
    bool ov = false;
    int namb = 0, ai = 0, aj = 0;
#pragma unroll 1
    for (int p0 = w0; p0 < nitems; p0 += T) {
        const int p = p0 + ln;
        bool near = false;
        int packed = 0;
        if (p < nitems) {
            const int d1 = div_nc(w, p); // offset - 1
            const int ci = p - d1 * nc, d = d1 + 1;
            int cj = ci + d;
            if (cj >= nc) cj -= nc;
            if (!(even && d == half && ci >= half)) {
                near = spheres_near(f, SM4[g4 + ci * nb + m], SM4[g4 + cj * nb + m], SMF[w.rad + ci] + SMF[w.rad + cj] + 1.001f);
                packed = (ci << 16) | cj;
                pairs += 1;
            }
        }
        const unsigned mask = __ballot_sync(FULL, near);
        if (near) lst[__popc(mask & ((1u << ln) - 1u))] = packed;
        __syncwarp();
        const int items = __popc(mask) * nb;
#pragma unroll 1
        for (int e = ln; e < items; e += 32) {
            const int pi = div_nb(w, e);
            const int pk = lst[pi];
            const int bi = (pk >> 16) * nb + (e - pi * nb), cj0 = (pk & 0xffff) * nb;
            const float4 s = SM4[g4 + bi];
#pragma unroll 1
            for (int b = 0; b < nb; ++b) {
                const float4 t = SM4[g4 + cj0 + b];
                const float r2 = r2_f32(f, s.x - t.x, s.y - t.y, s.z - t.z);
                ov |= (r2 < f.lo);
                if (!(r2 < f.lo) && !(r2 > f.hi)) {
                    if (namb == 0) { ai = bi; aj = cj0 + b; }
                    ++namb;
                }
            }
            pairs += (unsigned)nb;
        }
        // (the vote also orders this wave's reads of LST before the next wave's writes)
        if (T == 32) {
            if (__any_sync(FULL, ov)) return true;
        } else {
            if (__any_sync(FULL, ov)) *flag = 1;
            if (__any_sync(FULL, *flag != 0)) break;
        }
    }
    if (with_intra && w.npc > 0) { // same-chain pairs further apart than nexclude (shear / stretch, MCNS.py:198,247)
        const uchar2 *tab = W_ITAB(w);
        const int total = nc * w.npc;
#pragma unroll 1
        for (int e = w.lane; e < total; e += T) {
            const int c = e / w.npc;
            const uchar2 jb = tab[e - c * w.npc];
            const int bi = c * nb + jb.x, bj = c * nb + jb.y;
            const float4 s = SM4[g4 + bi], t = SM4[g4 + bj];
            const float r2 = r2_f32(f, s.x - t.x, s.y - t.y, s.z - t.z);
            ov |= (r2 < f.lo);
            if (!(r2 < f.lo) && !(r2 > f.hi)) {
                if (namb == 0) { ai = bi; aj = bj; }
                ++namb;
            }
            pairs += 1;
        }
    }
    if (gany<T>(ov)) return true;
    if (gany<T>(namb > 0)) {
        if (namb > 0)
            ov |= pair_r2(smem + sh, smem + sh + 9, smem[tp + 3 * aj] - smem[tp + 3 * ai], smem[tp + 3 * aj + 1] - smem[tp + 3 * ai + 1],
                          smem[tp + 3 * aj + 2] - smem[tp + 3 * ai + 2]) < 1.0;
        if (gany<T>(ov)) return true;
        if (gany<T>(namb > 1)) return trial_overlap_exact<T>(w, with_intra);
    }
    return false;
}

// all pairs of the TRIAL beads through the float32 filter: same enumeration, branch-free inner
// loop (8 independent pair tests per lane between votes); pairs inside the band are settled by
// one float64 pass afterwards, and only if nothing overlapped for certain.
template <int T>
__device__ __forceinline__ bool trial_overlap(const WCtx &w, bool with_intra, bool use_cells, unsigned &pairs)
{
    if (w.nc == 1) return trial_overlap_exact<T>(w, with_intra);
    if (use_cells) return trial_overlap_cells<T>(w, with_intra, pairs);
    if (w.nb > 1) return trial_overlap_pruned<T>(w, with_intra, pairs);
    const Filt f = load_filt(W_SF(w) + SF_TRIAL);
    const int nb = w.nb, N = w.N, half = N >> 1, g4 = W_G4(w);
    const bool even = (N & 1) == 0;
    const int RB = T < N ? T : N;
    const int DG = T < N ? 1 : T / N;
    const int li = w.lane % RB, dg = w.lane / RB;
    const bool lane_on = dg < DG;
    const int nex_eff = with_intra ? w.nex : 0x3fffffff; // same-chain pairs are tested iff d > nex_eff
    const int tp = W_TP(w), sh = W_SH(w) + 18;
    bool ov = false, overflow = false;
    constexpr int U = 2;
    for (int i0 = 0; i0 < N; i0 += RB) {
        const int i = i0 + li;
        const bool row_on = lane_on && i < N;
        const int ii = row_on ? i : 0;
        const float4 si = SM4[g4 + ii];
        const int dse = nb - 1 - (ii % nb);                                   // d <= dse: same chain
        const int dmax = row_on ? half - ((even && ii >= half) ? 1 : 0) : 0;  // offsets this lane owns
        int namb = 0, aj0 = 0, aj1 = 0; // pairs of this row that fell into the band (first two kept)
        for (int d0 = 1; d0 <= half; d0 += DG * U) { // (lane-uniform trip count: the vote below is a collective)
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int d = d0 + dg + u * DG;
                int j = ii + d;
                if (j >= N) j -= N;
                const bool valid = (d <= dmax) && ((d > dse) || (d > nex_eff));
                const float4 sj = SM4[g4 + (valid ? j : ii)];
                const float r2 = r2_f32(f, sj.x - si.x, sj.y - si.y, sj.z - si.z);
                ov |= valid && (r2 < f.lo);
                const bool amb = valid && !(r2 < f.lo) && !(r2 > f.hi);
                aj1 = (amb && namb == 1) ? j : aj1;
                aj0 = (amb && namb == 0) ? j : aj0;
                namb += amb;
                pairs += valid;
            }
            if (gany<T>(ov)) return true;
        }
        // pairs inside the band are settled in float64 right away (rigid chains can keep an
        // intra-chain pair inside the band for a whole walk, so this path must stay cheap)
        if (gany<T>(namb > 0)) {
            if (namb > 0)
                ov |= pair_r2(smem + sh, smem + sh + 9, smem[tp + 3 * aj0] - smem[tp + 3 * ii], smem[tp + 3 * aj0 + 1] - smem[tp + 3 * ii + 1],
                              smem[tp + 3 * aj0 + 2] - smem[tp + 3 * ii + 2]) < 1.0;
            if (namb > 1)
                ov |= pair_r2(smem + sh, smem + sh + 9, smem[tp + 3 * aj1] - smem[tp + 3 * ii], smem[tp + 3 * aj1 + 1] - smem[tp + 3 * ii + 1],
                              smem[tp + 3 * aj1 + 2] - smem[tp + 3 * ii + 2]) < 1.0;
            overflow |= namb > 2;
            if (gany<T>(ov)) return true;
        }
    }
    if (gany<T>(overflow)) return trial_overlap_exact<T>(w, with_intra);
    return false;
}

// chains follow a cell change rigidly (first bead keeps its fractional coordinates): beads at
// `src` under the inverse cell at smem[hi_old] -> beads at `dst` under the cell at smem[h_new], and
// their float32 shadows under its inverse at smem[hi_new].  Out of line, one copy: a box move applies
// it once, a rejected shear / stretch twice (instruction-cache footprint, profiles/r01_*).
template <int T>
__device__ __noinline__ void follow_cell(WCtx w, int src, int dst, int dst4, int hi_old, int h_new, int hi_new)
{
    double Hio[9], Hn[9], Hin[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) { Hio[k] = smem[hi_old + k]; Hn[k] = smem[h_new + k]; Hin[k] = smem[hi_new + k]; }
#pragma unroll 1
    for (int i = w.lane; i < w.N; i += T) {
        const int a = src + 3 * (div_nb(w, i) * w.nb);
        const double ax = smem[a], ay = smem[a + 1], az = smem[a + 2];
        double f0, f1, f2, n0, n1, n2;
        vecmat(ax, ay, az, Hio, f0, f1, f2);
        vecmat(f0, f1, f2, Hn, n0, n1, n2);
        const double x = smem[src + 3 * i] + (n0 - ax), y = smem[src + 3 * i + 1] + (n1 - ay), z = smem[src + 3 * i + 2] + (n2 - az);
        smem[dst + 3 * i] = x; smem[dst + 3 * i + 1] = y; smem[dst + 3 * i + 2] = z;
        SM4[dst4 + i] = shadow_of(x, y, z, Hin);
    }
}

template <int T> __device__ __forceinline__ void commit_trial(const WCtx &w)
{
    const int p = W_P(w), tp = W_TP(w), f4 = W_F4(w), g4 = W_G4(w), sh = W_SH(w), sf = W_SF(w);
#pragma unroll 1
    for (int i = w.lane; i < 3 * w.N; i += T) smem[p + i] = smem[tp + i];
#pragma unroll 1
    for (int i = w.lane; i < w.N; i += T) SM4[f4 + i] = SM4[g4 + i];
    if (w.lane < 18) smem[sh + w.lane] = smem[sh + 18 + w.lane];
    if (w.lane < 9) SMF[sf + w.lane] = SMF[sf + SF_TRIAL + w.lane];
    gsync<T>();
}

// delta_H of a shear / stretch proposal exactly as the reference's Python builds it: box_shear_step
// (MCNS.py:148-184: Gram-Schmidt of the two other cell vectors with np.dot, delta_H[k] = (H[k] + rv1 v1 + rv2 v2) - H[k])
// and box_stretch_step (MCNS.py:220-234: delta_H[i] = H[i] e^rv - H[i], delta_H[j] = H[j] e^-rv - H[j]).  dH must be zeroed.
__device__ __forceinline__ void shape_delta(const double *H, int type, int idx0, int idx1, double p0, double p1, double *dH)
{
    if (type == HANS_SHEAR) {
        const int k = idx0;
        const int o0 = (k == 0) ? 1 : 0;
        const int o1 = (k == 2) ? 1 : 2;
        double v1[3] = {H[3 * o0], H[3 * o0 + 1], H[3 * o0 + 2]};
        double v2[3] = {H[3 * o1], H[3 * o1 + 1], H[3 * o1 + 2]};
        const double n1 = sqrt(dot3_blas(v1, v1)); // np.dot, MCNS.py:161
        v1[0] /= n1; v1[1] /= n1; v1[2] /= n1;
        const double pr = dot3_blas(v1, v2);
        v2[0] -= v1[0] * pr; v2[1] -= v1[1] * pr; v2[2] -= v1[2] * pr;
        const double n2 = sqrt(dot3_blas(v2, v2));
        v2[0] /= n2; v2[1] /= n2; v2[2] /= n2;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double nw = H[3 * k + a] + (p0 * v1[a] + p1 * v2[a]);
            dH[3 * k + a] = nw - H[3 * k + a];
        }
    } else {
        const int i = idx0, j = idx1;
        const double e1 = hd_exp(p0), e2 = hd_exp(-p0);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            dH[3 * i + a] = H[3 * i + a] * e1 - H[3 * i + a];
            dH[3 * j + a] = H[3 * j + a] * e2 - H[3 * j + a];
        }
    }
}

struct DecP { int accepted, reason; double boltz; unsigned pairs; };

// alk.alkane_box_resize(pressure, ibox, 0) + acceptance (MCNS.py:318,349-357), and
// box_shear_step (MCNS.py:135-207) / box_stretch_step (MCNS.py:209-252) + accept or revert by -dH
// (MCNS.py:360-368).  The trial lives in TP, so a rejected volume move restores the saved cell and
// chains exactly for free.
template <int T>
__device__ __noinline__ DecP move_box(WCtx w, int q, double volume_limit, int mode)
{
    const PropRec *p = W_RING(w) + q;
    const hans_cfg *c = W_CFG(w);
    const int sh = W_SH(w);
    DecP out;
    out.accepted = 0; out.reason = HANS_REJ_INVALID; out.boltz = 0.0; out.pairs = 0;
    double H[9], Hn[9], Hin[9], dH[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) { H[k] = smem[sh + k]; dH[k] = 0.0; }
    const bool is_vol = p->type == HANS_VOL;
    double boltz_f = 1.0;
    if (is_vol) {
        const double vol = fabs(det3(H));
        const double vnew_t = vol + p->p[0];
        if (!(vnew_t > 0.0)) return out;
        // clearly above the ceiling: rejected before any further arithmetic (MCNS.py:351)
        if (mode == HANS_MODE_FUSED && (vnew_t - volume_limit) > 1e-9 * fabs(volume_limit)) {
            out.reason = HANS_REJ_CEILING;
            return out;
        }
        const double ratio = vnew_t / vol;
        const double s = hd_cbrt(ratio);
#pragma unroll
        for (int k = 0; k < 9; ++k) Hn[k] = H[k] * s;
        const double new_volume = fabs(det3(Hn)); // MCNS.py:350
        // exp(-P dV + nchains ln(V'/V)) (MCNS.py:266) as (V'/V)^nchains exp(-P dV)
        boltz_f = hd_powi(ratio, c->nchains);
        if (c->pressure != 0.0) boltz_f = boltz_f * hd_exp(-c->pressure * p->p[0]);
        out.boltz = boltz_f;
        if (mode == HANS_MODE_FUSED) {
            // conjunction of MCNS.py:351 evaluated cheapest-first; the O(N^2) overlap test is last
            if (!((new_volume - volume_limit) < 2.2204460492503131e-16)) { out.reason = HANS_REJ_CEILING; return out; }
            if (!(p->u_acc < boltz_f)) { out.reason = HANS_REJ_BOLTZ; return out; }
        }
    } else {
        shape_delta(H, p->type, p->idx0, p->idx1, p->p[0], p->p[1], dH);
#pragma unroll
        for (int k = 0; k < 9; ++k) Hn[k] = H[k] + dH[k];
    }
    hinv(Hn, Hin);
    if (w.lane == 0) {
#pragma unroll
        for (int k = 0; k < 9; ++k) { smem[sh + 18 + k] = Hn[k]; smem[sh + 27 + k] = Hin[k]; }
    }
    gsync<T>();
    follow_cell<T>(w, W_P(w), W_TP(w), W_G4(w), sh + 9, sh + 18, sh + 27); // alk.alkane_change_box, MCNS.py:186,238
    filter_params(sh + 18, W_SF(w) + SF_TRIAL, w.lane == 0);
    if (T > 32 && w.lane == 0) reinterpret_cast<int *>(smem)[w.lst - 3] = 0; // the early-exit flag of trial_overlap_pruned
    gsync<T>();
    int reason = HANS_ACCEPT;
    if (!is_vol) {
        if (aspect_fails(Hn, c->min_ar)) reason = HANS_REJ_ASPECT;                // MCNS.py:194,243
        else if (angle_fails(Hn, c->cos_min_ang)) reason = HANS_REJ_ANGLE;        // MCNS.py:196,245
    }
    // the cell list serves the box move when it is maintained anyway, or (bit 1) when bounding spheres prune nothing: the
    // typical reach of a pair of chains (two bounding radii of a zig-zag chain + sigma) exceeds half the smallest
    // perpendicular width of the trial cell, so the chain-pair pass would test all bead pairs -- and the trial configuration
    // is dense enough for the list to pay (measured crossover, profiles/r02_cells_probe.md)
    bool use_cells = (w.cells & 1) != 0;
    if (!use_cells && (w.cells & 2)) {
        const float reach = 2.0f * 0.82f * (float)c->bondlength * (float)(w.nb >> 1) + 1.0f;
        use_cells = !(reach < SMF[W_SF(w) + SF_TRIAL + 8]) && !((float)((double)w.N / fabs(det3(Hn))) < w.rho_box);
    }
    if (reason == HANS_ACCEPT && trial_overlap<T>(w, !is_vol, use_cells, out.pairs)) reason = HANS_REJ_OVERLAP; // MCNS.py:198,247
    if (mode == HANS_MODE_PROPOSE || reason == HANS_ACCEPT) {
        commit_trial<T>(w);
        if (w.cells & 1) {
            // (a shape rejection or a single-chain walker never built the trial list)
            if (reason == HANS_REJ_ASPECT || reason == HANS_REJ_ANGLE || w.nc == 1) cl_build<T>(w, sh + 9, W_F4(w), CL_DIMS(w), CL_HEAD(w), CL_NEXT(w));
            else cl_commit_trial<T>(w);
        }
    } else if (!is_vol) {
        // MCNS.py:366-368: the reference reverts by applying -dH through alkane_change_box, which
        // is not an exact restore; reproduce that arithmetic from the trial state.  (A rejected
        // volume move needs nothing: the saved cell and chains were never overwritten.)
        double H2[9], H2i[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) H2[k] = Hn[k] + (-1.0 * dH[k]);
        hinv(H2, H2i);
        gsync<T>(); // every thread has read the old cell
        if (w.lane == 0) {
#pragma unroll
            for (int k = 0; k < 9; ++k) { smem[sh + k] = H2[k]; smem[sh + 9 + k] = H2i[k]; }
        }
        gsync<T>();
        follow_cell<T>(w, W_TP(w), W_P(w), W_F4(w), sh + 27, sh, sh + 9);
        filter_params(sh, W_SF(w), w.lane == 0);
        gsync<T>();
        if (w.cells & 1) cl_build<T>(w, sh + 9, W_F4(w), CL_DIMS(w), CL_HEAD(w), CL_NEXT(w)); // the shadows were recomputed
    }
    out.accepted = reason == HANS_ACCEPT; out.reason = reason;
    out.boltz = reason == HANS_ACCEPT ? (is_vol ? boltz_f : 1.0) : 0.0;
    return out;
}

// ---------------------------------------------------------------------------------------------
// state load / store
// ---------------------------------------------------------------------------------------------
template <int T>
__device__ __forceinline__ void load_walker(const WCtx &w, const double *gH, const double *gR, int wi, bool shadows)
{
    const double *r = gR + (size_t)wi * w.N * 3;
    const int p = W_P(w), sh = W_SH(w);
#pragma unroll 1
    for (int i = w.lane; i < 3 * w.N; i += T) smem[p + i] = r[i];
    double H[9], Hi[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) H[k] = gH[(size_t)wi * 9 + k];
    hinv(H, Hi);
    if (w.lane == 0) {
#pragma unroll
        for (int k = 0; k < 9; ++k) { smem[sh + k] = H[k]; smem[sh + 9 + k] = Hi[k]; }
    }
    gsync<T>();
    if (shadows) {
        const int f4 = W_F4(w);
#pragma unroll 1
        for (int i = w.lane; i < w.N; i += T) SM4[f4 + i] = shadow_of(smem[p + 3 * i], smem[p + 3 * i + 1], smem[p + 3 * i + 2], smem + sh + 9);
        if (w.nb > 1)
#pragma unroll 1
            for (int c = w.lane; c < w.nc; c += T) SMF[w.rad + c] = chain_radius(p + 3 * c * w.nb, w.nb);
        filter_params(sh, W_SF(w), w.lane == 0);
        gsync<T>();
        if (w.cells & 1) cl_build<T>(w, sh + 9, f4, CL_DIMS(w), CL_HEAD(w), CL_NEXT(w));
    }
}

template <int T>
__device__ __forceinline__ void store_walker(const WCtx &w, double *gH, double *gR, int wi)
{
    double *r = gR + (size_t)wi * w.N * 3;
    const int p = W_P(w), sh = W_SH(w);
#pragma unroll 1
    for (int i = w.lane; i < 3 * w.N; i += T) r[i] = smem[p + i];
    if (w.lane < 9) gH[(size_t)wi * 9 + w.lane] = smem[sh + w.lane];
}

// ---------------------------------------------------------------------------------------------
// MC_run (MCNS.py:276-392) for a batch of walkers; REPLAY = proposals come from memory.
// Moves are processed in batches of T: first every lane produces the proposal of ONE move of the
// batch (Philox and the transcendental part of a proposal run once per move instead of once per
// lane), then the batch is applied move by move by the whole group.
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
    const hans_proposal *props; // replay
    hans_decision *out;
    int mode;
    int team_res; // warp-per-move kernel: int offset of its tail area in shared memory (set by the launcher)
    unsigned long long ctr_add; // proposals from memory: what the walk adds to the walkers' proposal counters
    // 0 = no cell list; 1 = the per-walker cell list serves every overlap test (walk_kernel only); 2 = only box moves
    // whose trial configuration is denser than rho_split (walk_kernel and walk_team_kernel)
    int cells_mode;
    double rho_split;
};

template <int T, bool REPLAY>
__global__ void __launch_bounds__(cta_threads<T>()) walk_kernel(const WalkArgs a)
{
    constexpr int GPC = cta_threads<T>() / T;
    WCtx w = setup_ctx<T>(a.cfg, a.cells_mode ? (int)cell_smem_bytes(a.cfg.nchains * a.cfg.nbeads, a.cells_mode == 2) : 0);
    w.rho_box = (float)a.rho_split;
    const int g = threadIdx.x / T;
    const double limit = a.d_limit ? *a.d_limit : a.limit;
    const int mode = a.mode;
    for (int k0 = blockIdx.x * GPC; k0 < a.n_active; k0 += gridDim.x * GPC) {
        const int k = k0 + g;
        if (k < a.n_active) {
            const int wi = a.ids ? a.ids[k] : k;
            w.cells = a.cells_mode; // 1: lists for every test; 2: box moves only, above rho_split
            load_walker<T>(w, a.H, a.R, wi, true);
            const unsigned long long ctr0 = REPLAY ? 0ull : a.ctr[wi];
            unsigned att[6] = {0, 0, 0, 0, 0, 0}, acc[6] = {0, 0, 0, 0, 0, 0}; // this lane's share
            unsigned pairs = 0;
            unsigned long long pairs_total = 0;
            PropRec *ring = W_RING(w);
            for (int m0 = 0; m0 < a.nmoves; m0 += T) {
                const int nbatch = a.nmoves - m0 < T ? a.nmoves - m0 : T;
                int my_type = -1;
                if (w.lane < nbatch) {
                    PropRec *slot = ring + w.lane;
                    if (REPLAY) {
                        const uint4 *q = reinterpret_cast<const uint4 *>(a.props + (size_t)k * a.nmoves + m0 + w.lane);
                        uint4 *d = reinterpret_cast<uint4 *>(slot);
                        d[0] = q[0]; d[1] = q[1]; d[2] = q[2]; d[3] = q[3];
                        my_type = slot->type;
                    } else {
                        const Prop pr = gen_proposal(W_CFG(w), a.seed, a.gid_base + (uint32_t)wi,
                                                     ctr0 + (unsigned long long)(m0 + w.lane));
                        slot->type = pr.type; slot->ichain = pr.ichain; slot->idx0 = pr.idx0; slot->idx1 = pr.idx1;
                        slot->p[0] = pr.p[0]; slot->p[1] = pr.p[1]; slot->p[2] = pr.p[2]; slot->p[3] = pr.p[3];
                        slot->u_acc = pr.u_acc;
                        my_type = pr.type;
                    }
                }
                gsync<T>();
                int my_acc = 0, my_reason = 0;
                double my_boltz = 0.0;
                for (int q = 0; q < nbatch; ++q) {
                    const int type = ring[q].type;
                    int accepted, reason;
                    double boltz;
                    if (type >= HANS_TRANS && type <= HANS_DIH) {
                        const Dec d = move_chain<T>(w, ring + q, mode, pairs);
                        accepted = d.accepted; reason = d.reason; boltz = d.boltz;
                    } else {
                        const DecP d = move_box<T>(w, q, limit, mode);
                        accepted = d.accepted; reason = d.reason; boltz = d.boltz; pairs += d.pairs;
                    }
                    if (w.lane == q) { my_acc = accepted; my_reason = reason; my_boltz = boltz; }
                }
                pairs_total += pairs;
                pairs = 0;
                if (my_type >= 0) {
#pragma unroll
                    for (int t = 0; t < 6; ++t) { att[t] += (my_type == t); acc[t] += (my_type == t && my_acc); }
                    if (REPLAY && a.out) {
                        hans_decision *o = a.out + (size_t)k * a.nmoves + m0 + w.lane;
                        o->accepted = my_acc; o->reason = my_reason; o->boltz = my_boltz;
                    }
                }
                gsync<T>(); // the ring is rewritten by the next batch
            }
            store_walker<T>(w, a.H, a.R, wi);
            if (a.attempted) { // reduce the per-lane counters over the group
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
                else if (a.ctr) a.ctr[wi] += a.ctr_add; // proposals came from hans_gen_proposals
                if (a.vol) a.vol[wi] = fabs(det3(smem + W_SH(w)));
            }
            if (a.pair_tests) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) pairs_total += __shfl_down_sync(FULL, pairs_total, o);
                if ((threadIdx.x & 31) == 0 && pairs_total) atomicAdd(a.pair_tests, pairs_total);
            }
            gsync<T>();
        }
        if (T != 32) __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// Proposals of a walk, generated ahead of it: one thread per (walker, move).  Proposals do not depend
// on the state, so this throughput-bound part (Philox, Box-Muller, quaternions) runs at full
// occupancy here instead of on the latency-bound critical path -- and outside the instruction-cache
// footprint -- of the walk kernels, which then read 64 bytes per move from HBM.
// props[k][m] = proposal number ctr[walker k] + ctr_off + m of walker k (same stream as the fused kernels).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gen_kernel(const hans_cfg cfg, const unsigned long long *ctr, unsigned long long ctr_off,
                                                  const int32_t *ids, int n_active, int nmoves, unsigned long long seed,
                                                  uint32_t gid_base, hans_proposal *props)
{
    const size_t total = (size_t)n_active * (size_t)nmoves;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(t / (size_t)nmoves), m = (int)(t - (size_t)k * nmoves);
        const int wi = ids ? ids[k] : k;
        const Prop pr = gen_proposal(&cfg, seed, gid_base + (uint32_t)wi, ctr[wi] + ctr_off + (unsigned long long)m);
        PropRec r;
        r.type = pr.type; r.ichain = pr.ichain; r.idx0 = pr.idx0; r.idx1 = pr.idx1;
        r.p[0] = pr.p[0]; r.p[1] = pr.p[1]; r.p[2] = pr.p[2]; r.p[3] = pr.p[3];
        r.u_acc = pr.u_acc; r.pad = 0.0;
        const uint4 *src = reinterpret_cast<const uint4 *>(&r);
        uint4 *dst = reinterpret_cast<uint4 *>(props + t);
        dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
    }
}

// ---------------------------------------------------------------------------------------------
// the non-hot kernels (predicate, change_box, growth): float64 only
// ---------------------------------------------------------------------------------------------
// alk.alkane_check_chain_overlap (MCNS.py:198,247,543) for a batch of walkers
template <int T>
__global__ void __launch_bounds__(cta_threads<T>())
overlap_kernel(const hans_cfg cfg, const double *gH, const double *gR, const int32_t *ids, int n, int inter_only,
               int32_t *out)
{
    constexpr int GPC = cta_threads<T>() / T;
    const WCtx w = setup_ctx<T>(cfg);
    const int g = threadIdx.x / T;
    for (int k0 = blockIdx.x * GPC; k0 < n; k0 += gridDim.x * GPC) {
        const int k = k0 + g;
        if (k < n) {
            const int wi = ids ? ids[k] : k;
            load_walker<T>(w, gH, gR, wi, false);
            double H[9], Hi[9];
#pragma unroll
            for (int q = 0; q < 9; ++q) { H[q] = smem[W_SH(w) + q]; Hi[q] = smem[W_SH(w) + 9 + q]; }
            const bool ov = all_overlap64<T>(w, W_P(w), H, Hi, !inter_only);
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
    constexpr int GPC = cta_threads<T>() / T;
    const WCtx w = setup_ctx<T>(cfg);
    const int g = threadIdx.x / T;
    for (int k0 = blockIdx.x * GPC; k0 < n; k0 += gridDim.x * GPC) {
        const int k = k0 + g;
        if (k < n) {
            const int wi = ids ? ids[k] : k;
            load_walker<T>(w, gH, gR, wi, false);
            const int sh = W_SH(w);
            double Hn[9], Hin[9];
#pragma unroll
            for (int q = 0; q < 9; ++q) Hn[q] = smem[sh + q] + gdH[(size_t)k * 9 + q];
            hinv(Hn, Hin);
            if (w.lane == 0) {
#pragma unroll
                for (int q = 0; q < 9; ++q) { smem[sh + 18 + q] = Hn[q]; smem[sh + 27 + q] = Hin[q]; }
            }
            gsync<T>();
            follow_cell<T>(w, W_P(w), W_TP(w), W_G4(w), sh + 9, sh + 18, sh + 27);
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
            uint32_t gid_base, int chain_begin, int chain_end, uint32_t attempt_begin, int max_attempts, int32_t *status)
{
    constexpr int GPC = cta_threads<T>() / T;
    const WCtx w = setup_ctx<T>(cfg);
    const int g = threadIdx.x / T;
    const int nb = w.nb, P = W_P(w), cn = W_CN(w);
    for (int k0 = blockIdx.x * GPC; k0 < n; k0 += gridDim.x * GPC) {
        const int k = k0 + g;
        if (k < n) {
            const int wi = ids ? ids[k] : k;
            const uint32_t gid = gid_base + (uint32_t)wi;
            load_walker<T>(w, gH, gR, wi, false);
            double H[9], Hi[9];
#pragma unroll
            for (int q = 0; q < 9; ++q) { H[q] = smem[W_SH(w) + q]; Hi[q] = smem[W_SH(w) + 9 + q]; }
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
                    if (w.lane == 0) { smem[cn] = x0; smem[cn + 1] = y0; smem[cn + 2] = z0; }
                    if (nb > 1) {
                        const double z = 2.0 * ud - 1.0;
                        const double st2 = 1.0 - z * z;
                        const double st = sqrt(st2 > 0.0 ? st2 : 0.0);
                        double s, co;
                        hd_sincos(HD_TWO_PI * ue, &s, &co);
                        pc[0] = x0 + cfg.bondlength * (st * co);
                        pc[1] = y0 + cfg.bondlength * (st * s);
                        pc[2] = z0 + cfg.bondlength * z;
                        if (w.lane == 0) { smem[cn + 3] = pc[0]; smem[cn + 4] = pc[1]; smem[cn + 5] = pc[2]; }
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
                        if (w.lane == 0) { smem[cn + 3 * kb] = nw[0]; smem[cn + 3 * kb + 1] = nw[1]; smem[cn + 3 * kb + 2] = nw[2]; }
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
                            const double xj = smem[P + 3 * j], yj = smem[P + 3 * j + 1], zj = smem[P + 3 * j + 2];
                            for (int b = 0; b < nb; ++b)
                                ov |= (pair_r2(H, Hi, xj - smem[cn + 3 * b], yj - smem[cn + 3 * b + 1], zj - smem[cn + 3 * b + 2]) < 1.0);
                        }
                        if (gany<T>(ov)) break;
                    }
                    ov = gany<T>(ov);
                    if (!ov) {
                        const uchar2 *tab = W_ITAB(w);
                        for (int e = w.lane; e < w.npc; e += T) {
                            const uchar2 jb = tab[e];
                            ov |= pair_r2(H, Hi, smem[cn + 3 * jb.y] - smem[cn + 3 * jb.x], smem[cn + 3 * jb.y + 1] - smem[cn + 3 * jb.x + 1],
                                          smem[cn + 3 * jb.y + 2] - smem[cn + 3 * jb.x + 2]) < 1.0;
                        }
                        ov = gany<T>(ov);
                    }
                    if (!ov) {
                        for (int i = w.lane; i < 3 * nb; i += T) smem[P + 3 * ic * nb + i] = smem[cn + i];
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

} // namespace hb
