// hans_kernels.cu -- sm_100a kernels + C ABI of libhans_b200.so (see include/hans_b200.h).
//
// The walker kernels (MC_run batch, overlap predicate, change_box, growth) are in hans_walk.cuh;
// this file holds the small nested-sampling / cell kernels and the extern "C" entry points.
//
// Compile: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -lineinfo  (see
// __graft_entry__.build()).  -fmad=false is REQUIRED for bit-exact parity with the CPU checker.
#include "hans_lean.cuh"
#include "hans_team.cuh"

#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>

namespace hb {

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

// delta_H of n shear / stretch proposals on the cells of walkers ids[k] (box_shear_step MCNS.py:148-184,
// box_stretch_step MCNS.py:220-234); other proposal types give zero
__global__ void shape_delta_kernel(const double *gH, const int32_t *ids, int n, const hans_proposal *props, double *dH)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int wi = ids ? ids[k] : k;
    double H[9], d[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) { H[q] = gH[(size_t)wi * 9 + q]; d[q] = 0.0; }
    const hans_proposal p = props[k];
    const bool shear = p.type == HANS_SHEAR && (unsigned)p.idx0 < 3u;
    const bool stretch = p.type == HANS_STRETCH && (unsigned)p.idx0 < 3u && (unsigned)p.idx1 < 3u && p.idx0 != p.idx1;
    if (shear || stretch) shape_delta(H, p.type, p.idx0, p.idx1, p.p[0], p.p[1], d);
#pragma unroll
    for (int q = 0; q < 9; ++q) dH[(size_t)k * 9 + q] = d[q];
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

// local MAXLOC (mpihans.py:211-212): ties -> lowest index
__global__ void argmax_kernel(const double *vol, int n, double *out2)
{
    double best; int bi;
    block_argmax(vol, n, best, bi);
    if (threadIdx.x == 0) { out2[0] = best; out2[1] = (double)bi; }
}

// local MAXLOC (mpihans.py:211-212) and staging of the clone source (mpihans.py:226-229)
__device__ __forceinline__ void ns_pack_body(const hans_ns_args &a)
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
__global__ void ns_pack_kernel(const hans_ns_args a) { ns_pack_body(a); }

// global MAXLOC (mpihans.py:215), ceiling, log, trajectory staging, clone (mpihans.py:232-237),
// active walkers (mpihans.py:236-239)
__device__ __forceinline__ void ns_resolve_body(const hans_ns_args &a, const double *gathered)
{
    __shared__ double s_vmax;
    __shared__ int s_rank, s_idx;
    const int rec = 12 + 3 * a.N, n3 = 3 * a.N;
    const long long it = *a.d_iter;
    if (threadIdx.x == 0) {
        double best = -1.7976931348623157e308; int br = 0, bi = 0;
        for (int r = 0; r < a.world; ++r) {
            const double v = gathered[(size_t)r * rec];
            if (v > best) { best = v; br = r; bi = (int)gathered[(size_t)r * rec + 1]; }
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
        const double *sr = gathered + (size_t)src_rank * rec;
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
__global__ void ns_resolve_kernel(const hans_ns_args a) { ns_resolve_body(a, a.d_gathered); }

// ---------------------------------------------------------------------------------------------
// The exchange of an iteration over NVLink peer memory instead of a collective library: ONE kernel
// per rank does pack -> push -> wait -> resolve.  Every rank owns a buffer that all ranks of the node
// have mapped (CUDA IPC): flags[2][world] (u64, padded to 128 bytes) + records[2][world][12+3N]
// (float64); the leading index is the parity of the iteration.  A rank stores the live part of its
// record into slot [parity][rank] of EVERY rank's buffer (plain stores through the peer mapping,
// fence.sys), then raises flag [parity][rank] there to iteration + 1 (st.release.sys); it then spins
// (ld.acquire.sys) until the `world` flags of that parity in its OWN buffer have reached
// iteration + 1 and resolves from its own copy.  Two parities suffice: a rank can only write
// iteration i + 2 after it has seen every rank's flag of iteration i + 1, which each rank raises
// after it has finished reading iteration i (stream order).  Flags only grow, so nothing is reset.
// ---------------------------------------------------------------------------------------------
struct PeerTable { unsigned long long *base[HANS_PEER_MAX]; };
__host__ __device__ inline size_t peer_header_words(int world) { return (size_t)((2 * world + 15) / 16) * 16; }

__global__ void __launch_bounds__(512) ns_exchange_peer_kernel(const hans_ns_args a, const PeerTable pt, long long spin_limit)
{
    const int rec = 12 + 3 * a.N, world = a.world;
    const long long it = *a.d_iter;
    const int par = (int)(it & 1);
    const size_t hdr = peer_header_words(world);
    ns_pack_body(a); // -> a.d_rec
    __syncthreads();
    const unsigned long long K = (unsigned long long)a.nlocal * (unsigned long long)world;
    const int src_rank = (int)((ns_draw(a.seed, it, 0u, 0u) % K) / (unsigned long long)a.nlocal);
    const int len = src_rank == a.rank ? rec : 2; // only the owner of the clone source ships a configuration
    for (int p = 0; p < world; ++p) {
        double *dst = reinterpret_cast<double *>(pt.base[p] + hdr) + ((size_t)par * world + a.rank) * rec;
        for (int i = threadIdx.x; i < len; i += blockDim.x) dst[i] = a.d_rec[i];
    }
    __threadfence_system();
    __syncthreads();
    const unsigned long long want = (unsigned long long)it + 1ull;
    __shared__ int s_expired;
    if (threadIdx.x == 0) s_expired = 0;
    __syncthreads();
    if ((int)threadIdx.x < world) {
        unsigned long long *f = pt.base[threadIdx.x] + (size_t)par * world + a.rank;
        asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(f), "l"(want) : "memory");
        const unsigned long long *g = pt.base[a.rank] + (size_t)par * world + threadIdx.x;
        unsigned long long v;
        const long long t0 = clock64();
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(g) : "memory");
            // a peer that never raises its flag (a dead rank) must not hang this GPU for ever: the wait is bounded
            // (hans_ns_peer_timeout; default 60 s), an expired wait is counted and the iteration is NOT resolved
            if (spin_limit > 0 && v < want && clock64() - t0 > spin_limit) { atomicAdd(&g_diag[2], 1ull); s_expired = 1; break; }
        } while (v < want);
    }
    __syncthreads();
    if (s_expired) return; // stale records: nothing is cloned, the iteration counter does not advance; the host sees diag[2]
    ns_resolve_body(a, reinterpret_cast<const double *>(pt.base[a.rank] + hdr) + (size_t)par * world * rec);
}

// Order parameters of n configurations (nematic.py:83-137 nematic P2; nematic.py:30-55 translational), one warp
// per configuration, lanes over chains.
//   P2    largest eigenvalue of Q = mean_chains (3 u u^T - I) / 2, u = unit minimum-image end-to-end vector
//         (bead 0 -> bead nbeads-1; nematic.py:112-121).  Single beads have no director: P2 = 0.
//   trans |sum_chains exp(2 pi i (s_x + s_y + s_z))| / nchains over the fractional "centres" the reference uses:
//         the mean of beads 0 .. nbeads-2 (the slice a0:a1 of nematic.py:66-72 leaves the last bead out), or the
//         bead itself for nbeads = 1; every bead wrapped into the cell first, as ase's get_scaled_positions does.
__global__ void order_kernel(const hans_cfg cfg, const double *gH, const double *gR, int n, double *p2, double *trans)
{
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (wid >= n) return;
    const int nb = cfg.nbeads, nc = cfg.nchains;
    double H[9], Hi[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) H[k] = gH[(size_t)wid * 9 + k];
    hinv(H, Hi);
    const double *R = gR + (size_t)wid * nb * nc * 3;
    double q[6] = {0, 0, 0, 0, 0, 0}, re = 0.0, im = 0.0;
    for (int c = lane; c < nc; c += 32) {
        const double *r0 = R + (size_t)c * nb * 3;
        if (nb > 1) {
            const double *r1 = r0 + (size_t)(nb - 1) * 3;
            double s0, s1, s2, ux, uy, uz;
            vecmat(r1[0] - r0[0], r1[1] - r0[1], r1[2] - r0[2], Hi, s0, s1, s2);
            s0 -= hd_rint(s0); s1 -= hd_rint(s1); s2 -= hd_rint(s2);
            vecmat(s0, s1, s2, H, ux, uy, uz);
            const double inv = 1.0 / sqrt((ux * ux + uy * uy) + uz * uz);
            ux *= inv; uy *= inv; uz *= inv;
            q[0] += 1.5 * ux * ux - 0.5; q[1] += 1.5 * uy * uy - 0.5; q[2] += 1.5 * uz * uz - 0.5;
            q[3] += 1.5 * ux * uy; q[4] += 1.5 * ux * uz; q[5] += 1.5 * uy * uz;
        }
        // ase's get_scaled_positions() wraps every atom into [0, 1) before the mean is taken (nematic.py:66-72)
        double f0 = 0.0, f1 = 0.0, f2 = 0.0;
        const int m = nb > 1 ? nb - 1 : 1;
        for (int b = 0; b < m; ++b) {
            double g0, g1, g2;
            vecmat(r0[3 * b], r0[3 * b + 1], r0[3 * b + 2], Hi, g0, g1, g2);
            f0 += g0 - floor(g0); f1 += g1 - floor(g1); f2 += g2 - floor(g2);
        }
        f0 /= (double)m; f1 /= (double)m; f2 /= (double)m;
        double sn, cs;
        sincos(HD_TWO_PI * ((f0 + f1) + f2), &sn, &cs);
        re += cs; im += sn;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int k = 0; k < 6; ++k) q[k] += __shfl_down_sync(FULL, q[k], o);
        re += __shfl_down_sync(FULL, re, o);
        im += __shfl_down_sync(FULL, im, o);
    }
    if (lane != 0) return;
    if (trans) trans[wid] = sqrt(re * re + im * im) / (double)nc;
    if (p2) {
        double e = 0.0;
        if (nb > 1) {
            const double a11 = q[0] / nc, a22 = q[1] / nc, a33 = q[2] / nc, a12 = q[3] / nc, a13 = q[4] / nc, a23 = q[5] / nc;
            const double p1 = (a12 * a12 + a13 * a13) + a23 * a23;
            const double tq = ((a11 + a22) + a33) / 3.0;
            const double pp = ((a11 - tq) * (a11 - tq) + (a22 - tq) * (a22 - tq)) + ((a33 - tq) * (a33 - tq) + 2.0 * p1);
            const double pr = sqrt(pp / 6.0);
            if (pr < 1e-300) e = tq;
            else {
                const double b11 = (a11 - tq) / pr, b22 = (a22 - tq) / pr, b33 = (a33 - tq) / pr;
                const double b12 = a12 / pr, b13 = a13 / pr, b23 = a23 / pr;
                double r = 0.5 * ((b11 * (b22 * b33 - b23 * b23) - b12 * (b12 * b33 - b23 * b13)) + b13 * (b12 * b23 - b22 * b13));
                r = r < -1.0 ? -1.0 : (r > 1.0 ? 1.0 : r);
                e = tq + 2.0 * pr * cos(acos(r) / 3.0); // largest eigenvalue of a symmetric 3x3 matrix
            }
        }
        p2[wid] = e;
    }
}

// FP32 FFMA issue-rate probe (roofline denominator): 16 independent accumulators, 256 FFMAs per loop trip, so the
// loop overhead (compare + branch + counter) is ~1 % of the issue slots
#define HANS_FFMA_PER_TRIP 256
__global__ void __launch_bounds__(256) ffma_kernel(int iters, float *sink)
{
    float a[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) a[k] = threadIdx.x * 1e-3f + (float)k;
    const float m = 0.999f, c = 1e-3f;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < HANS_FFMA_PER_TRIP / 16; ++u)
#pragma unroll
            for (int k = 0; k < 16; ++k) a[k] = __fmaf_rn(a[k], m, c);
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += a[k];
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

static size_t cta_smem_bytes(const hans_cfg *c, int T, bool cells = false)
{
    const int gpc = (T == 32) ? 4 : 1;
    const int N = c->nchains * c->nbeads;
    return (walker_smem_bytes(N, c->nbeads, T) + (cells ? cell_smem_bytes(N) : 0)) * gpc + cta_extra_bytes(c->nbeads, c->nexclude);
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

constexpr size_t SMEM_MAX = 227 * 1024;

template <typename K> static int prep_kernel(K kernel, size_t smem)
{
    if (smem > SMEM_MAX) return fail(HANS_ELIMIT, "walker does not fit in shared memory for this kernel family");
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return HANS_OK;
}

// Walkers beyond 227 KB: the same kernels with the walker state in a per-CTA region of GLOBAL memory (namespace hbg,
// hans_walk_body.inc).  The regions are a stream-ordered allocation that lives for the launch.
struct BigState {
    double *base = nullptr;
    unsigned long long stride = 0; // doubles per CTA
    int grid = 0;
};
static int big_alloc(BigState &b, size_t bytes_per_cta, int ctas_wanted, cudaStream_t st)
{
    const int cap = sm_count() * 4;
    b.grid = ctas_wanted < cap ? (ctas_wanted > 0 ? ctas_wanted : 1) : cap;
    b.stride = (bytes_per_cta + 255) / 256 * 32; // doubles, 256-byte aligned regions
    CK(cudaMallocAsync((void **)&b.base, (size_t)b.grid * b.stride * sizeof(double), st));
    return HANS_OK;
}
static void big_free(BigState &b, cudaStream_t st) { if (b.base) cudaFreeAsync(b.base, st); b.base = nullptr; }

static int grid_for(int n, int T)
{
    const int gpc = (T == 32) ? 4 : 1;
    const int ctas = (n + gpc - 1) / gpc;
    const int cap = sm_count() * 16;
    return ctas < cap ? (ctas > 0 ? ctas : 1) : cap;
}

static int pick_T(const hans_cfg *c, int T)
{
    T &= ~HANS_TPW_CELLS; // (the flag only steers the walk kernels)
    if (T >= 0 && T < 32) T = hans_default_threads_per_walker(c); // 1..31 only steer the walk kernels
    if (T != 32 && T != 64 && T != 128 && T != 256) return -1;
    return T;
}

static int spec_grid(int n) { const int cap = sm_count() * 32; return n < cap ? n : cap; }

// Shapes the lean warp-per-walker kernel (hans_lean.cuh) runs: beads-per-chain counts it is
// instantiated for; single spheres up to 128 beads (every pair is tested), chains of any size that
// fits shared memory (the chain-level pruning leaves a warp's worth of pair tests per move).
static bool lean_applies(const hans_cfg *c)
{
    const int nb = c->nbeads, N = c->nchains * c->nbeads;
    if (!(nb == 1 || nb == 2 || nb == 4 || nb == 6 || nb == 8 || nb == 12)) return false;
    if (N <= 128) return true;
    // measured (profiles/r01_*), M moves/s: 32 chains x 12 beads (N = 384) 151 here vs 169 with one CTA of 64
    // threads per walker vs 110 with 128; 128 chains x 6 beads (N = 768) 99 here vs 137 with 128 threads:
    // the warp-per-walker kernel keeps the walkers below 256 beads
    if (nb == 1 || c->nchains > 64 || N >= 256) return false;
    return walker_smem_bytes(N, nb, 32) + cta_extra_bytes(nb, c->nexclude) <= 200 * 1024;
}

#define DISPATCH_NB(NB, ...)                                 \
    switch (NB) {                                            \
    case 1: { constexpr int NN = 1; __VA_ARGS__; } break;    \
    case 2: { constexpr int NN = 2; __VA_ARGS__; } break;    \
    case 4: { constexpr int NN = 4; __VA_ARGS__; } break;    \
    case 6: { constexpr int NN = 6; __VA_ARGS__; } break;    \
    case 8: { constexpr int NN = 8; __VA_ARGS__; } break;    \
    default: { constexpr int NN = 12; __VA_ARGS__; } break;  \
    }

// W = 1: one move at a time; 4 / 8: speculative windows (hans_lean.cuh)
template <bool REPLAY> static int launch_lean(const WalkArgs &a, int W, cudaStream_t st)
{
    const hans_cfg *c = &a.cfg;
    const size_t smem = walker_smem_bytes(c->nchains * c->nbeads, c->nbeads, 32) + window_smem_bytes(c->nbeads, W > 1 ? W : 0) +
                        cta_extra_bytes(c->nbeads, c->nexclude);
    int rc = HANS_OK;
#define LEAN_LAUNCH(WW, MB)                                                                   \
    DISPATCH_NB(c->nbeads, {                                                                  \
        rc = prep_kernel(walk_lean_kernel<NN, WW, REPLAY, MB>, smem);                         \
        if (rc) return rc;                                                                    \
        walk_lean_kernel<NN, WW, REPLAY, MB><<<spec_grid(a.n_active), 32, smem, st>>>(a);     \
    })
    // Register allocation by the size of the batch (one-warp CTAs; A/B builds on B200, M moves/s).
    // Uncapped (214 - 243 registers: 8 or 9 walkers per SM) is fastest per walker and, for chains, also in total at any batch
    // size but one: a batch of more than that and up to 12 walkers per SM leaves a short second wave, and the 168-register build (12 per SM)
    // takes it at once (C3, 1 536 walkers: 608 vs 490).  Single spheres lose less to the cap: beyond what the uncapped build holds the
    // 128-register build (16 per SM) wins (C2, 2 048 / 4 096 walkers: 1 876 / 2 046 vs 1 727 / 1 800).
    const int per_sm = (a.n_active + sm_count() - 1) / sm_count();
    int fit = 8; // one-warp CTAs of the uncapped build an SM holds (registers, shared memory): asked of the runtime
    if (W == 1) {
        DISPATCH_NB(c->nbeads, { if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&fit, walk_lean_kernel<NN, 1, REPLAY, 1>, 32, smem) != cudaSuccess || fit < 1) { (void)cudaGetLastError(); fit = 8; } })
    }
    const int mb = per_sm <= fit ? 1 : c->nbeads == 1 ? 16 : per_sm <= 12 ? 12 : 1;
    if (W == 1) { if (mb == 16) { LEAN_LAUNCH(1, 16); } else if (mb == 12) { LEAN_LAUNCH(1, 12); } else { LEAN_LAUNCH(1, 1); } }
    else { LEAN_LAUNCH(8, 1); }
#undef LEAN_LAUNCH
    CK(cudaGetLastError());
    return HANS_OK;
}

// window size of the lean kernel for a request: 0 = default, 1 / 4 / 8 explicit; -1 = not applicable
static int pick_W(const hans_cfg *c, int tpw)
{
    if (tpw >= 32 || !lean_applies(c)) return -1;
    if (tpw == 0) return hans_default_window(c);
    if (tpw == 8 && c->nchains * c->nbeads > 128) return -1; // the windows' bit words cover 4 x 32 beads
    return (tpw == 1 || tpw == 8) ? tpw : -1;
}

// One CTA of NWARP warps per walker, one warp per move of a speculative window (hans_team.cuh): chains only
// (the pruning works on chain bounding spheres), and the walker plus NWARP trial chains must fit shared memory.
static size_t team_smem_bytes(const hans_cfg *c, int nwarp)
{
    return walker_smem_bytes(c->nchains * c->nbeads, c->nbeads, 32 * nwarp) + window_smem_bytes(c->nbeads, nwarp) +
           team_tail_bytes(nwarp) + cta_extra_bytes(c->nbeads, c->nexclude);
}
static bool team_applies(const hans_cfg *c, int nwarp)
{
    return c->nbeads > 1 && c->nchains >= 2 && team_smem_bytes(c, nwarp) <= 227 * 1024;
}

template <bool REPLAY> static int launch_team(WalkArgs a, int nwarp, cudaStream_t st)
{
    const hans_cfg *c = &a.cfg;
    a.team_res = (int)((walker_smem_bytes(c->nchains * c->nbeads, c->nbeads, 32 * nwarp) + window_smem_bytes(c->nbeads, nwarp)) / 4);
    const size_t smem = team_smem_bytes(c, nwarp);
    const int cap = sm_count() * 8, grid = a.n_active < cap ? a.n_active : cap;
    int rc = HANS_OK;
#define TEAM_LAUNCH(NN, WW)                                                        \
    {                                                                              \
        rc = prep_kernel(walk_team_kernel<NN, WW, REPLAY>, smem);                  \
        if (rc) return rc;                                                         \
        walk_team_kernel<NN, WW, REPLAY><<<grid, 32 * WW, smem, st>>>(a);          \
    }
#define TEAM_NB(WW)                                                                \
    switch (c->nbeads) {                                                           \
    case 6: TEAM_LAUNCH(6, WW); break;                                             \
    case 12: TEAM_LAUNCH(12, WW); break;                                           \
    default: TEAM_LAUNCH(0, WW); break;                                            \
    }
    if (nwarp == 4) { TEAM_NB(4); }
    else { TEAM_NB(8); }
#undef TEAM_NB
#undef TEAM_LAUNCH
    CK(cudaGetLastError());
    return HANS_OK;
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
    if (N <= 512) return 64;
    return 128; // (256 threads per walker measured slower at every BASELINE shape: barriers dominate)
}

int hans_default_window(const hans_cfg *cfg)
{
    if (!cfg || !lean_applies(cfg)) return 0;
    // Measured on B200 (profiles/r01_*): with the proposals generated ahead and the code footprint
    // inside the instruction cache, one move at a time beats the speculative windows at every
    // BASELINE shape -- phase B costs as many instructions per accepted move as a serial move.
    // The windows stay available through threads_per_walker = 8.
    return 1;
}

int hans_default_team(const hans_cfg *cfg)
{
    if (!cfg || check_cfg(cfg) != HANS_OK) return 0;
    // Measured on B200 (profiles/r01_probe_team_kernel.jsonl), M moves/s: 128 chains x 6 beads 148 (CTA of 128 threads
    // sharing every move) vs 265 / 343 with 4 / 8 warps, one move each; 32 chains x 12 beads 169 (CTA of 64) vs 199 / 140
    // (windows of distinct chains are short with 32 chains); 16 x 6 beads: the warp-per-walker kernel stays 2x ahead.
    if (cfg->nbeads < 2 || cfg->nchains * cfg->nbeads < 256) return 0;
    if (cfg->nchains >= 64 && team_applies(cfg, 8)) return 8;
    if (cfg->nchains >= 24 && team_applies(cfg, 4)) return 4;
    return 0;
}

// threads per walker of the cell-list kernel for a shape
static int cells_T(const hans_cfg *c)
{
    static int forced = -1;
    if (forced < 0) { const char *e = getenv("HANS_CELLS_T"); forced = e ? atoi(e) : 0; }
    if (forced == 32 || forced == 64 || forced == 128 || forced == 256) return forced;
    const int N = c->nchains * c->nbeads;
    return N <= 128 ? 32 : (N <= 1024 ? 128 : 256);
}

// the cooperative kernel (all threads of a walker's group share every move), optionally with the per-walker cell list
static int launch_coop(const hans_cfg *cfg, const WalkArgs &a, bool from_memory, int T, cudaStream_t st)
{
    int rc = HANS_OK;
    const size_t smem = cta_smem_bytes(cfg, T, a.cells_mode != 0);
    if (smem > SMEM_MAX) { // state in global memory
        BigState b;
        rc = big_alloc(b, smem, (a.n_active + (T == 32 ? 3 : 0)) / (T == 32 ? 4 : 1), st);
        if (rc) return rc;
        hbg::WalkArgs g;
        static_assert(sizeof(hbg::WalkArgs) == sizeof(WalkArgs), "the two state spaces share the argument layout");
        memcpy(&g, &a, sizeof g);
        g.state_base = b.base; g.state_stride = b.stride;
        if (from_memory) { DISPATCH_T(T, { hbg::walk_kernel<TT, true, true><<<b.grid, cta_threads<TT>(), 16, st>>>(g); }); }
        else { DISPATCH_T(T, { hbg::walk_kernel<TT, false, true><<<b.grid, cta_threads<TT>(), 16, st>>>(g); }); }
        const cudaError_t e = cudaGetLastError();
        big_free(b, st);
        if (e != cudaSuccess) return cuda_fail(e, "walk_kernel (state in global memory)");
        return HANS_OK;
    }
#define COOP_LAUNCH(RP, CLV)                                                                                  \
    DISPATCH_T(T, {                                                                                           \
        rc = prep_kernel(walk_kernel<TT, RP, CLV>, smem);                                                     \
        if (rc) return rc;                                                                                    \
        walk_kernel<TT, RP, CLV><<<grid_for(a.n_active, TT), cta_threads<TT>(), smem, st>>>(a);               \
    })
    if (a.cells_mode) { if (from_memory) { COOP_LAUNCH(true, true); } else { COOP_LAUNCH(false, true); } }
    else { if (from_memory) { COOP_LAUNCH(true, false); } else { COOP_LAUNCH(false, false); } }
#undef COOP_LAUNCH
    CK(cudaGetLastError());
    return HANS_OK;
}

static int launch_walk(const hans_cfg *cfg, WalkArgs &a, bool from_memory, int threads_per_walker, cudaStream_t st)
{
    const bool want_cells = (threads_per_walker & HANS_TPW_CELLS) != 0;
    threads_per_walker &= ~HANS_TPW_CELLS;
    if (want_cells) { // explicit: the cooperative kernel with the cell list for every walker
        const int T = threads_per_walker == 0 ? cells_T(cfg) : pick_T(cfg, threads_per_walker);
        if (T < 0 || threads_per_walker == 1 || threads_per_walker == 8 || threads_per_walker == 24 || threads_per_walker == 28)
            return fail(HANS_EINVAL, "HANS_TPW_CELLS goes with threads_per_walker 0, 32, 64, 128 or 256");
        a.cells_mode = 1;
        return launch_coop(cfg, a, from_memory, T, st);
    }
    const bool dflt = threads_per_walker == 0;
    if (dflt && a.mode == HANS_MODE_FUSED && !lean_applies(cfg) && hans_default_team(cfg))
        threads_per_walker = 20 + hans_default_team(cfg);
    if ((threads_per_walker == 24 || threads_per_walker == 28) && a.mode == HANS_MODE_FUSED) {
        const int nwarp = threads_per_walker - 20;
        if (!team_applies(cfg, nwarp))
            return fail(HANS_ELIMIT, "the warp-per-move kernel (threads_per_walker 24, 28) needs chains (nbeads > 1) that fit shared memory");
        return from_memory ? launch_team<true>(a, nwarp, st) : launch_team<false>(a, nwarp, st);
    }
    if (threads_per_walker < 24 && a.mode == HANS_MODE_FUSED) { // PROPOSE leaves trial states: cooperative kernel
        const int W = pick_W(cfg, threads_per_walker);
        if (W > 0) return from_memory ? launch_lean<true>(a, W, st) : launch_lean<false>(a, W, st);
        if (threads_per_walker != 0)
            return fail(HANS_ELIMIT, "the warp-per-walker kernel (threads_per_walker 1, 8) needs nchains*nbeads <= 128 and nbeads in {1,2,4,6,8,12}");
    }
    const int T = pick_T(cfg, threads_per_walker);
    if (T < 0) return fail(HANS_EINVAL, "threads_per_walker must be 0, 1, 8, 24, 28, 32, 64, 128 or 256 (optionally | HANS_TPW_CELLS)");
    if (dflt && cta_smem_bytes(cfg, T) > SMEM_MAX) {
        // a walker that does not fit shared memory: state in global memory, and the per-walker cell list for every overlap
        // test (all pairs of thousands of beads is what the list is for)
        a.cells_mode = 1;
        return launch_coop(cfg, a, from_memory, 256, st);
    }
    return launch_coop(cfg, a, from_memory, T, st);
}

static int gen_launch(const hans_cfg *cfg, const uint64_t *d_ctr, uint64_t ctr_off, const int32_t *d_ids, int n_active,
                      int nmoves, uint64_t seed, uint32_t gid_base, hans_proposal *d_props, cudaStream_t st)
{
    const size_t total = (size_t)n_active * (size_t)nmoves;
    const size_t want = (total + 255) / 256, cap = (size_t)sm_count() * 64;
    gen_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(*cfg, (const unsigned long long *)d_ctr, ctr_off, d_ids,
                                                                   n_active, nmoves, seed, gid_base, d_props);
    CK(cudaGetLastError());
    return HANS_OK;
}

int hans_gen_proposals(const hans_cfg *cfg, const uint64_t *d_ctr, const int32_t *d_ids, int n_active, int nmoves,
                       uint64_t seed, uint32_t gid_base, hans_proposal *d_props, void *stream)
{
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if (!d_ctr || !d_props) return fail(HANS_EINVAL, "d_ctr and d_props are required");
    if (n_active < 0 || nmoves < 0) return fail(HANS_EINVAL, "negative count");
    if (n_active == 0 || nmoves == 0) return HANS_OK;
    return gen_launch(cfg, d_ctr, 0, d_ids, n_active, nmoves, seed, gid_base, d_props, (cudaStream_t)stream);
}

// How hans_mc_run_batch_ws cuts a walk of `total` moves: the workspace holds the whole walk -> one generate + one
// walk; it holds less -> chunks of what fits, generated and walked in turn through the one buffer; no workspace ->
// generation inside the walk kernel.  (Measured and dropped: cutting a walk that fits into four chunks generated on a
// side stream while the previous chunk is walked hides the generator -- 5 % of a C2 step -- but every chunk reloads its
// walker and has its own tail of late CTAs: C2 1 465 -> 1 294, C3 559 -> 520, C5 373 -> 364 M moves/s.)
struct WsPlan { int chunk; int nchunks; };
static WsPlan ws_plan(int n_active, int total, bool have_ws, uint64_t workspace_bytes)
{
    WsPlan p = {0, 0};
    if (!have_ws || n_active <= 0 || total <= 0) return p;
    const uint64_t per_move = (uint64_t)n_active * sizeof(hans_proposal);
    const uint64_t chunk = workspace_bytes / per_move;
    p.chunk = chunk >= (uint64_t)total ? total : (int)(chunk & ~(uint64_t)31); // a multiple of 32: the kernels take proposals 32 at a time
    p.nchunks = p.chunk ? (total + p.chunk - 1) / p.chunk : 0;
    return p;
}

int hans_mc_run_launches(int n_active, int nmoves, int have_workspace, uint64_t workspace_bytes)
{
    if (n_active <= 0 || nmoves <= 0) return 0; // nothing is launched for an empty batch
    const WsPlan p = ws_plan(n_active, nmoves, have_workspace != 0, workspace_bytes);
    return p.nchunks ? 2 * p.nchunks : 1;
}

int hans_mc_run_batch_ws(const hans_cfg *cfg, double *d_H, double *d_R, uint64_t *d_ctr, const int32_t *d_ids,
                         int n_active, int sweeps, const double *d_volume_limit, double volume_limit, uint64_t seed,
                         uint32_t gid_base, uint64_t *d_attempted, uint64_t *d_accepted, double *d_vol,
                         uint64_t *d_pair_tests, int threads_per_walker, void *d_workspace, uint64_t workspace_bytes,
                         void *stream)
{
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if (!d_H || !d_R || !d_ctr) return fail(HANS_EINVAL, "d_H, d_R and d_ctr are required");
    if (n_active < 0 || sweeps < 0) return fail(HANS_EINVAL, "negative n_active / sweeps");
    if ((d_attempted == nullptr) != (d_accepted == nullptr)) return fail(HANS_EINVAL, "d_attempted and d_accepted go together");
    if (n_active == 0) return HANS_OK;
    WalkArgs a;
    memset(&a, 0, sizeof a);
    a.cfg = *cfg; a.H = d_H; a.R = d_R; a.ctr = (unsigned long long *)d_ctr; a.ids = d_ids; a.n_active = n_active;
    a.nmoves = sweeps * hans_moves_per_sweep(cfg->nbeads, cfg->nchains);
    a.d_limit = d_volume_limit; a.limit = volume_limit; a.seed = seed; a.gid_base = gid_base;
    a.attempted = (unsigned long long *)d_attempted; a.accepted = (unsigned long long *)d_accepted;
    a.vol = d_vol; a.pair_tests = (unsigned long long *)d_pair_tests; a.mode = HANS_MODE_FUSED;
    cudaStream_t st = (cudaStream_t)stream;
    const int total = a.nmoves;
    const WsPlan plan = ws_plan(n_active, total, d_workspace != nullptr, workspace_bytes);
    if (plan.nchunks == 0) return launch_walk(cfg, a, false, threads_per_walker, st); // no workspace: generate in the kernel
    hans_proposal *ws = (hans_proposal *)d_workspace;
    const int chunk = plan.chunk;
    for (int m0 = 0; m0 < total; m0 += chunk) {
        const int n = total - m0 < chunk ? total - m0 : chunk;
        rc = gen_launch(cfg, d_ctr, 0, d_ids, n_active, n, seed, gid_base, ws, st);
        if (rc) return rc;
        a.nmoves = n;
        a.props = ws;
        a.out = nullptr;
        a.ctr_add = (unsigned long long)n;
        rc = launch_walk(cfg, a, true, threads_per_walker, st); // advances d_ctr by n
        if (rc) return rc;
    }
    return HANS_OK;
}

int hans_mc_run_batch(const hans_cfg *cfg, double *d_H, double *d_R, uint64_t *d_ctr, const int32_t *d_ids,
                      int n_active, int sweeps, const double *d_volume_limit, double volume_limit, uint64_t seed,
                      uint32_t gid_base, uint64_t *d_attempted, uint64_t *d_accepted, double *d_vol,
                      uint64_t *d_pair_tests, int threads_per_walker, void *stream)
{
    return hans_mc_run_batch_ws(cfg, d_H, d_R, d_ctr, d_ids, n_active, sweeps, d_volume_limit, volume_limit, seed, gid_base,
                                d_attempted, d_accepted, d_vol, d_pair_tests, threads_per_walker, nullptr, 0, stream);
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
    WalkArgs a;
    memset(&a, 0, sizeof a);
    a.cfg = *cfg; a.H = d_H; a.R = d_R; a.ids = d_ids; a.n_active = n_active; a.nmoves = nmoves;
    a.limit = volume_limit; a.props = d_props; a.out = d_out; a.mode = mode; a.ctr_add = (unsigned long long)nmoves;
    return launch_walk(cfg, a, true, threads_per_walker, (cudaStream_t)stream);
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
    if (smem > SMEM_MAX) {
        BigState b;
        rc = big_alloc(b, smem, (n + (T == 32 ? 3 : 0)) / (T == 32 ? 4 : 1), st);
        if (rc) return rc;
        DISPATCH_T(T, { hbg::overlap_kernel<TT><<<b.grid, cta_threads<TT>(), 16, st>>>(*cfg, d_H, d_R, d_ids, n, inter_only, d_out, b.base, b.stride); });
        big_free(b, st);
        CK(cudaGetLastError());
        return HANS_OK;
    }
    DISPATCH_T(T, {
        rc = prep_kernel(overlap_kernel<TT>, smem);
        if (rc) return rc;
        overlap_kernel<TT><<<grid_for(n, TT), cta_threads<TT>(), smem, st>>>(*cfg, d_H, d_R, d_ids, n, inter_only, d_out, nullptr, 0);
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
    if (smem > SMEM_MAX) {
        BigState b;
        rc = big_alloc(b, smem, (n + (T == 32 ? 3 : 0)) / (T == 32 ? 4 : 1), st);
        if (rc) return rc;
        DISPATCH_T(T, { hbg::change_box_kernel<TT><<<b.grid, cta_threads<TT>(), 16, st>>>(*cfg, d_H, d_R, d_ids, n, d_dH, b.base, b.stride); });
        big_free(b, st);
        CK(cudaGetLastError());
        return HANS_OK;
    }
    DISPATCH_T(T, {
        rc = prep_kernel(change_box_kernel<TT>, smem);
        if (rc) return rc;
        change_box_kernel<TT><<<grid_for(n, TT), cta_threads<TT>(), smem, st>>>(*cfg, d_H, d_R, d_ids, n, d_dH, nullptr, 0);
    });
    CK(cudaGetLastError());
    return HANS_OK;
}

int hans_shape_delta(const double *d_H, const int32_t *d_ids, int n, const hans_proposal *d_props, double *d_dH, void *stream)
{
    if (!d_H || !d_props || !d_dH) return fail(HANS_EINVAL, "NULL device pointer");
    if (n < 0) return fail(HANS_EINVAL, "negative count");
    if (n == 0) return HANS_OK;
    shape_delta_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(d_H, d_ids, n, d_props, d_dH);
    CK(cudaGetLastError());
    return HANS_OK;
}

static int launch_grow(const hans_cfg *cfg, const double *d_H, double *d_R, const int32_t *d_ids, int n, uint64_t seed,
                       uint32_t gid_base, int cb, int ce, uint32_t attempt0, int max_attempts, int32_t *d_status,
                       void *stream)
{
    const int T = pick_T(cfg, 0);
    const size_t smem = cta_smem_bytes(cfg, T);
    cudaStream_t st = (cudaStream_t)stream;
    int rc = HANS_OK;
    if (smem > SMEM_MAX) {
        BigState b;
        rc = big_alloc(b, smem, (n + (T == 32 ? 3 : 0)) / (T == 32 ? 4 : 1), st);
        if (rc) return rc;
        DISPATCH_T(T, { hbg::grow_kernel<TT><<<b.grid, cta_threads<TT>(), 16, st>>>(*cfg, d_H, d_R, d_ids, n, seed, gid_base, cb, ce, attempt0,
                                                                               max_attempts, d_status, b.base, b.stride); });
        big_free(b, st);
        CK(cudaGetLastError());
        return HANS_OK;
    }
    DISPATCH_T(T, {
        rc = prep_kernel(grow_kernel<TT>, smem);
        if (rc) return rc;
        grow_kernel<TT><<<grid_for(n, TT), cta_threads<TT>(), smem, st>>>(*cfg, d_H, d_R, d_ids, n, seed, gid_base, cb, ce,
                                                                          attempt0, max_attempts, d_status, nullptr, 0);
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
    return launch_grow(cfg, d_H, d_R, nullptr, n, seed, gid_base, 0, -1, 0u, max_attempts, d_status, stream);
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
    return launch_grow(cfg, d_H + off_h, d_R + off_r, nullptr, 1, seed, gid, ichain, ichain + 1, attempt, 1, d_ifail,
                       stream);
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

static long long g_peer_spin_limit = (long long)(60.0 * 1.9e9); // SM clock cycles (default 60 s); 0 = no limit
int hans_ns_peer_timeout(double seconds)
{
    g_peer_spin_limit = seconds > 0.0 ? (long long)(seconds * 1.9e9) : 0;
    return HANS_OK;
}

uint64_t hans_ns_peer_bytes(int N, int world)
{
    if (N < 1 || world < 1) return 0;
    return 8ull * (peer_header_words(world) + 2ull * (uint64_t)world * (uint64_t)(12 + 3 * N));
}

int hans_peer_alloc(uint64_t bytes, void **d_ptr, unsigned char *handle64)
{
    if (!d_ptr || !handle64 || bytes == 0) return fail(HANS_EINVAL, "bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void *p = nullptr;
    CK(cudaMalloc(&p, bytes));
    CK(cudaMemset(p, 0, bytes));
    CK(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return cuda_fail(e, "cudaIpcGetMemHandle"); }
    memcpy(handle64, &h, 64);
    *d_ptr = p;
    return HANS_OK;
}

int hans_peer_open(const unsigned char *handle64, void **d_ptr)
{
    if (!d_ptr || !handle64) return fail(HANS_EINVAL, "bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    CK(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return HANS_OK;
}

int hans_peer_close(void *d_ptr) { if (d_ptr) CK(cudaIpcCloseMemHandle(d_ptr)); return HANS_OK; }
int hans_peer_free(void *d_ptr) { if (d_ptr) CK(cudaFree(d_ptr)); return HANS_OK; }

int hans_ns_exchange_peer(const hans_ns_args *a, void *const *peer_bufs, void *stream)
{
    int rc = check_ns(a);
    if (rc) return rc;
    if (!peer_bufs) return fail(HANS_EINVAL, "peer_bufs is NULL");
    if (a->world > HANS_PEER_MAX) return fail(HANS_ELIMIT, "more ranks than HANS_PEER_MAX");
    PeerTable pt;
    memset(&pt, 0, sizeof pt);
    for (int r = 0; r < a->world; ++r) {
        if (!peer_bufs[r]) return fail(HANS_EINVAL, "NULL peer buffer");
        pt.base[r] = (unsigned long long *)peer_bufs[r];
    }
    ns_exchange_peer_kernel<<<1, 512, 0, (cudaStream_t)stream>>>(*a, pt, g_peer_spin_limit);
    CK(cudaGetLastError());
    return HANS_OK;
}

int hans_order_params(const hans_cfg *cfg, const double *d_H, const double *d_R, int n, double *d_p2, double *d_trans,
                      void *stream)
{
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if (!d_H || !d_R) return fail(HANS_EINVAL, "NULL device pointer");
    if (n < 0) return fail(HANS_EINVAL, "negative count");
    if (n == 0) return HANS_OK;
    order_kernel<<<(n + 3) / 4, 128, 0, (cudaStream_t)stream>>>(*cfg, d_H, d_R, n, d_p2, d_trans);
    CK(cudaGetLastError());
    return HANS_OK;
}

int hans_diag_counters(uint64_t *out4, int reset)
{
    if (!out4) return fail(HANS_EINVAL, "out4 is NULL");
    unsigned long long h[4] = {0, 0, 0, 0};
    CK(cudaMemcpyFromSymbol(h, g_diag, sizeof h));
    for (int i = 0; i < 4; ++i) out4[i] = h[i];
    if (reset) {
        unsigned long long z[4] = {0, 0, 0, 0};
        CK(cudaMemcpyToSymbol(g_diag, z, sizeof z));
    }
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
    float best = 0.f;
    for (int rep = 0; rep < 3; ++rep) { // best of three: the probe is a peak, not an average
        CK(cudaEventRecord(e0, st));
        ffma_kernel<<<blocks, threads, 0, st>>>(iters, sink);
        CK(cudaEventRecord(e1, st));
        CK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep == 0 || ms < best) best = ms;
    }
    *tflops = 2.0 * (double)HANS_FFMA_PER_TRIP * (double)iters * blocks * threads / (best * 1e-3) / 1e12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    return HANS_OK;
}

} // extern "C"
