// ARCHIVED EXPERIMENT -- not part of the library build.  Measured slower than hans_lean.cuh (DESIGN.md section 3,
// profiles/r01_ncu_c2_two_warps_per_walker_experiment.txt).  To rebuild it, WCtx needs redirectable trial-chain
// fields (`cn`, `c4`, used by W_CN / W_C4) and hans_kernels.cu a launch_pair dispatch; see the git history.
//
// hans_pair.cuh -- MC_run with TWO WARPS per walker: two consecutive single-chain moves evaluated at
// the same time, settled in order.
//
// Why: at ~1 000 walkers per GPU the walk is latency-bound -- throughput is
// (number of independent warps) / (serial instructions per move x ~4.4 cycles), a lone warp issues no
// faster than one that shares its scheduler, and the SMs have room for twice the warps
// (profiles/r01_*).  So the serial stream of a walker is cut in two: for a pair
// of consecutive single-chain moves (q, q+1), warp 0 evaluates move q and warp 1 evaluates move q+1
// against the SAME state, concurrently.  Warp 1 keeps the part of its result that concerns the chain
// c_q move q would move apart from the rest.  Then, in order:
//   move q    accepted iff nothing overlaps; warp 0 commits;
//   move q+1  if q was rejected: its result stands.  If q was committed: its overlaps with the old
//             position of c_q are void and warp 1 re-tests its trial chain against the new c_q only;
//             if q+1 moves c_q itself (its trial was built from the old position) or had a pair inside
//             the float32 band, warp 1 re-evaluates it from scratch.
// Every decision is taken against exactly the state the serial order produces, so trajectories are
// bit-identical to the other kernels and to the CPU checker (tests/test_gpu_parity.py).  Moves that
// are not plain single-chain moves (box moves, malformed proposals) are executed by warp 0 alone
// with the routines of hans_lean.cuh / hans_walk.cuh while warp 1 waits.
//
// Reference call sites replaced: NesSa/MCNS.py:304-383 (the move loop of MC_run).
#pragma once

#include "hans_lean.cuh"

namespace hb {

// result of evaluating a trial chain against the current state, apart for one chain `cx`
struct Eval {
    bool ov;    // certainly overlaps a chain other than cx
    bool ovx;   // certainly overlaps chain cx
    bool doubt; // a pair inside the filter's band and no certain overlap to make it irrelevant
};

__device__ __forceinline__ float warp_min(float v)
{
    // (r^2 may round to a tiny negative value: as integers all negatives sort below all positives, and
    // any negative result is "< lo" like the true minimum)
    return __int_as_float(__reduce_min_sync(FULL, __float_as_int(v)));
}

// single sphere, translation: trial bead + shadow in registers
__device__ __forceinline__ Eval eval_sphere(const WCtx &w, const CellRegs &cr, const PropRec *p, int ichain, int cx, /* cx by value */
                                            double &x, double &y, double &z, float4 &cs, unsigned &pairs)
{
    const int P = W_P(w) + 3 * ichain, f4 = W_F4(w), N = w.N;
    x = smem[P] + p->p[0]; y = smem[P + 1] + p->p[1]; z = smem[P + 2] + p->p[2];
    cs = shadow_of(x, y, z, cr.Hi);
    const Filt &f = cr.f;
    float rmin = 3.0e38f, rminx = 3.0e38f;
    if (cx == ichain) cx = -1; // (the chain itself is never a partner)
#pragma unroll 2
    for (int j = w.lane; j < N; j += 32) {
        const float4 s = SM4[f4 + j];
        const float r2 = r2_f32(f, s.x - cs.x, s.y - cs.y, s.z - cs.z);
        if (j == cx) rminx = r2;
        else if (j != ichain) rmin = fminf(rmin, r2);
    }
    pairs += (unsigned)((N - 1 - w.lane + 32) >> 5);
    const float m = warp_min(rmin), mx = warp_min(rminx);
    Eval e;
    e.ov = m < f.lo; e.ovx = mx < f.lo;
    e.doubt = (!e.ov && !(m > f.hi)) || (!e.ovx && !(mx > f.hi));
    return e;
}

__device__ __forceinline__ void commit_sphere(const WCtx &w, int ichain, double x, double y, double z, float4 cs)
{
    if (w.lane == 0) {
        const int P = W_P(w) + 3 * ichain;
        smem[P] = x; smem[P + 1] = y; smem[P + 2] = z;
        SM4[W_F4(w) + ichain] = cs;
    }
}

// chains: trial chain into the warp's own CN / C4, chain-level pruning, overlaps with chain cx apart
template <int NB>
__device__ __forceinline__ Eval eval_chain(const WCtx &w, const CellRegs &cr, const PropRec *p, int ichain, int cx,
                                           int &b0, bool &intra, float &Rt, unsigned &pairs)
{
    const int nc = w.nc, m = NB >> 1, f4 = W_F4(w), cn = W_CN(w), c4 = W_C4(w);
    const int *lst = reinterpret_cast<const int *>(smem) + w.lst;
    const Filt &f = cr.f;
    make_trial(w, p, w.lane, 32, cn, c4, b0, intra);
    __syncwarp();
    Rt = intra ? chain_radius(cn, NB) : SMF[w.rad + ichain];
    float4 tr[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) tr[b] = SM4[c4 + b];
    bool ov = false, amb = false, ovx = false, ambx = false;
    for (int cb = 0; cb < nc; cb += 32) {
        const int c = cb + w.lane;
        bool near = false;
        if (c < nc && c != ichain) near = spheres_near(f, SM4[f4 + c * NB + m], tr[m], Rt + SMF[w.rad + c] + 1.001f);
        const int items = near_list<32>(w, near, c) * NB;
        for (int e = w.lane; e < items; e += 32) {
            const int ci = e / NB, ch = lst[ci];
            const float4 s = SM4[f4 + ch * NB + (e - ci * NB)];
            bool o = false, a = false;
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                const float r2 = r2_f32(f, s.x - tr[b].x, s.y - tr[b].y, s.z - tr[b].z);
                const bool moved = b >= b0;
                o |= moved && (r2 < f.lo);
                a |= moved && !(r2 > f.hi);
            }
            if (ch == cx) { ovx |= o; ambx |= a; } else { ov |= o; amb |= a; }
            pairs += (unsigned)(NB - b0);
        }
        pairs += (c < nc) ? 1u : 0u;
        if (cb + 32 < nc) __syncwarp();
    }
    if (intra) { // dihedral: the chain's own pairs beyond the exclusion count as "not cx"
        const uchar2 *tab = W_ITAB(w);
        for (int e = w.lane; e < w.npc; e += 32) {
            const uchar2 jb = tab[e];
            if ((int)jb.y >= b0) {
                const float4 pa = SM4[c4 + jb.x], pb = SM4[c4 + jb.y];
                const float r2 = r2_f32(f, pb.x - pa.x, pb.y - pa.y, pb.z - pa.z);
                ov |= (r2 < f.lo);
                amb |= !(r2 > f.hi);
                pairs += 1;
            }
        }
    }
    const unsigned code = __reduce_or_sync(FULL, (ov ? 1u : 0u) | (ovx ? 2u : 0u) | (amb ? 4u : 0u) | (ambx ? 8u : 0u));
    Eval e;
    e.ov = (code & 1u) != 0u; e.ovx = (code & 2u) != 0u;
    e.doubt = (!e.ov && (code & 4u)) || (!e.ovx && (code & 8u));
    return e;
}

template <int NB> __device__ __forceinline__ void commit_chain(const WCtx &w, int ichain, bool intra, float Rt)
{
    const int P = W_P(w) + 3 * NB * ichain, f4 = W_F4(w) + NB * ichain, cn = W_CN(w), c4 = W_C4(w);
    for (int k = w.lane; k < 3 * NB; k += 32) smem[P + k] = smem[cn + k];
    for (int k = w.lane; k < NB; k += 32) SM4[f4 + k] = SM4[c4 + k];
    if (intra && w.lane == 0) SMF[w.rad + ichain] = Rt;
}

// my trial chain (b >= b0) against the COMMITTED chain cx only
template <int NB>
__device__ __forceinline__ Eval refit_chain(const WCtx &w, const CellRegs &cr, int cx, int b0, float Rt, unsigned &pairs)
{
    const int m = NB >> 1, f4 = W_F4(w) + NB * cx, c4 = W_C4(w);
    const Filt &f = cr.f;
    bool o = false, a = false;
    if (spheres_near(f, SM4[f4 + m], SM4[c4 + m], Rt + SMF[w.rad + cx] + 1.001f) && w.lane < NB) {
        const float4 s = SM4[f4 + w.lane];
        for (int b = b0; b < NB; ++b) {
            const float4 t = SM4[c4 + b];
            const float r2 = r2_f32(f, s.x - t.x, s.y - t.y, s.z - t.z);
            o |= (r2 < f.lo);
            a |= !(r2 > f.hi);
        }
        pairs += (unsigned)(NB - b0);
    }
    const unsigned code = __reduce_or_sync(FULL, (o ? 1u : 0u) | (a ? 2u : 0u));
    Eval e;
    e.ov = false; e.ovx = (code & 1u) != 0u; e.doubt = !e.ovx && (code & 2u);
    return e;
}

// one move from scratch by one warp (decide + commit): the serial routines of hans_lean.cuh
template <int NB>
__device__ __forceinline__ int serial_plain(const WCtx &w, const CellRegs &cr, const PropRec *p, int ichain, unsigned &pairs)
{
    if (NB == 1) return lean_translate_sphere(w, cr, p, ichain, pairs);
    return lean_move_chain<(NB > 1 ? NB : 2)>(w, cr, p, ichain, pairs);
}

#ifndef HANS_PAIR_MIN_CTAS
#define HANS_PAIR_MIN_CTAS 7
#endif
template <int NB, bool REPLAY>
__global__ void __launch_bounds__(64, HANS_PAIR_MIN_CTAS) walk_pair_kernel(const WalkArgs a)
{
    constexpr int T = 64;
    const int role = threadIdx.x >> 5;
    // one walker per CTA; the layout of a 64-thread group (ring and near list of 64 entries, one extra
    // trial chain behind the ring), both warps see it with lane = 0..31
    WCtx w = setup_ctx_g<64, 1>(a.cfg, 1);
    w.lane = threadIdx.x & 31;
    if (role == 1) { w.cn = w.wc; w.c4 = w.w4; w.lst += 32; } // warp 1: its own trial chain and near list
    int *flag = reinterpret_cast<int *>(smem) + (w.lst - (role ? 32 : 0)) - 4; // 4 ints in front of LST
    const double limit = a.d_limit ? *a.d_limit : a.limit;
    for (int k = blockIdx.x; k < a.n_active; k += gridDim.x) {
        const int wi = a.ids ? a.ids[k] : k;
        {   // both warps load (64 threads)
            WCtx wl = w;
            wl.lane = threadIdx.x;
            load_walker<T>(wl, a.H, a.R, wi, true);
        }
        CellRegs cr = load_cell_regs(w);
        const unsigned long long ctr0 = REPLAY ? 0ull : a.ctr[wi];
        unsigned att[6] = {0, 0, 0, 0, 0, 0}, acc[6] = {0, 0, 0, 0, 0, 0}; // this thread's moves (thread t: move t of a batch)
        unsigned pairs = 0;
        unsigned long long pairs_total = 0;
        PropRec *ring = W_RING(w);
        for (int m0 = 0; m0 < a.nmoves; m0 += T) {
            const int nbatch = a.nmoves - m0 < T ? a.nmoves - m0 : T;
            int my_type = -1;
            bool my_plain = false;
            if ((int)threadIdx.x < nbatch) {
                PropRec *slot = ring + threadIdx.x;
                if (REPLAY) {
                    const uint4 *q = reinterpret_cast<const uint4 *>(a.props + (size_t)k * a.nmoves + m0 + threadIdx.x);
                    uint4 *d = reinterpret_cast<uint4 *>(slot);
                    d[0] = q[0]; d[1] = q[1]; d[2] = q[2]; d[3] = q[3];
                    my_type = slot->type;
                } else {
                    const Prop pr = gen_proposal(W_CFG(w), a.seed, a.gid_base + (uint32_t)wi,
                                                 ctr0 + (unsigned long long)(m0 + threadIdx.x));
                    slot->type = pr.type; slot->ichain = pr.ichain; slot->idx0 = pr.idx0; slot->idx1 = pr.idx1;
                    slot->p[0] = pr.p[0]; slot->p[1] = pr.p[1]; slot->p[2] = pr.p[2]; slot->p[3] = pr.p[3];
                    slot->u_acc = pr.u_acc;
                    my_type = pr.type;
                }
                my_plain = (my_type == HANS_TRANS || (NB > 1 && my_type == HANS_ROT) ||
                            (my_type == HANS_DIH && NB >= 4 && slot->idx0 >= 3 && slot->idx0 <= NB - 1)) &&
                           (unsigned)slot->ichain < (unsigned)w.nc;
            }
            const unsigned pl = __ballot_sync(FULL, my_plain);
            if (w.lane == 0) flag[2 + role] = (int)pl;
            __syncthreads();
            const unsigned long long plain = (unsigned long long)(unsigned)flag[2] | ((unsigned long long)(unsigned)flag[3] << 32);
            unsigned long long accmask = 0; // accepted moves of the batch (CTA-uniform)
            int q = 0;
            // the two result words alternate between flag[0..1] and flag[2..3] from step to step, so a step
            // never rewrites words the other warp may still be reading; the first step of a batch uses
            // flag[0..1] (flag[2..3] still hold the plain mask until everybody is past its first barrier)
            int step = 0;
            while (q < nbatch) {
                int *fl = flag + 2 * (step & 1);
                ++step;
                const bool p0 = (plain >> q) & 1ull, p1 = (q + 1 < nbatch) && ((plain >> (q + 1)) & 1ull);
                if (p0 && p1) {
                    // ---- a pair of plain single-chain moves: both evaluated now, settled in order ----------
                    const PropRec *pq = ring + q, *pm = ring + q + role; // move q, my move
                    const int c0 = pq->ichain, ichain = pm->ichain;
                    const int cx = role ? c0 : -1;
                    double x = 0.0, y = 0.0, z = 0.0;
                    float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
                    int b0 = 0;
                    bool intra = false;
                    float Rt = 0.0f;
                    Eval e;
                    const bool same = role && ichain == c0; // my trial is built from a chain move q may move
                    if (NB == 1) e = eval_sphere(w, cr, pm, ichain, cx, x, y, z, cs, pairs);
                    else e = eval_chain<(NB > 1 ? NB : 2)>(w, cr, pm, ichain, cx, b0, intra, Rt, pairs);
                    bool blocked0 = e.ov; // (role 0: cx = -1, everything is in ov)
                    if (role == 0 && e.doubt) { // warp 0 sees the true state: float64 settles it at once
                        if (NB == 1 && w.lane == 0) { const int cn = W_CN(w); smem[cn] = x; smem[cn + 1] = y; smem[cn + 2] = z; }
                        __syncwarp();
                        blocked0 = lean_recheck(w, ichain, b0, true, intra);
                    }
                    if (role == 0 && w.lane == 0) fl[0] = blocked0 ? 1 : 0;
                    __syncthreads();                                                        // B1
                    const bool acc0 = fl[0] == 0;
                    const bool commit0 = acc0 && (pq->u_acc < 1.0); // MCNS.py:372 with boltz = 1
                    if (commit0) {
                        if (role == 0) {
                            if (NB == 1) commit_sphere(w, ichain, x, y, z, cs);
                            else commit_chain<(NB > 1 ? NB : 2)>(w, ichain, intra, Rt);
                        }
                        __syncthreads();                                                    // B2
                    }
                    if (role == 1) {
                        int ok;
                        if ((commit0 && same) || e.doubt) {
                            ok = serial_plain<NB>(w, cr, pm, ichain, pairs); // from scratch on the settled state (commits)
                        } else {
                            bool blocked = e.ov;
                            bool redo = false;
                            if (!blocked) {
                                if (!commit0) blocked = e.ovx;
                                else if (NB == 1) {
                                    const float4 s = SM4[W_F4(w) + c0];
                                    const float r2 = r2_f32(cr.f, s.x - cs.x, s.y - cs.y, s.z - cs.z);
                                    blocked = r2 < cr.f.lo;
                                    redo = !blocked && !(r2 > cr.f.hi);
                                    pairs += (w.lane == 0) ? 1u : 0u;
                                } else {
                                    const Eval r = refit_chain<(NB > 1 ? NB : 2)>(w, cr, c0, b0, Rt, pairs);
                                    blocked = r.ovx;
                                    redo = r.doubt;
                                }
                            }
                            if (redo) {
                                ok = serial_plain<NB>(w, cr, pm, ichain, pairs);
                            } else {
                                ok = blocked ? 0 : 1;
                                if (ok && (pm->u_acc < 1.0)) {
                                    if (NB == 1) commit_sphere(w, ichain, x, y, z, cs);
                                    else commit_chain<(NB > 1 ? NB : 2)>(w, ichain, intra, Rt);
                                }
                            }
                        }
                        if (w.lane == 0) fl[1] = ok;
                    }
                    __syncthreads();                                                        // B3
                    accmask |= (unsigned long long)(acc0 ? 1u : 0u) << q;
                    accmask |= (unsigned long long)(fl[1] ? 1u : 0u) << (q + 1);
                    q += 2;
                    continue;
                }
                // ---- a single move by warp 0 (a plain move without a partner, or a box move / malformed one) --
                if (role == 0) {
                    const PropRec *p = ring + q;
                    const int type = p->type, ichain = p->ichain;
                    int ok, cell = 0;
                    if (p0) {
                        ok = serial_plain<NB>(w, cr, p, ichain, pairs);
                    } else {
                        DecP d;
                        if (type >= HANS_TRANS && type <= HANS_DIH) {
                            d = lean_move_chain_generic(w, q);
                        } else {
                            d = move_box<32>(w, q, limit, HANS_MODE_FUSED);
                            cell = 1;
                        }
                        pairs += d.pairs;
                        ok = d.accepted ? 1 : 0;
                        if (w.lane == 0) { ring[q].idx0 = d.reason; ring[q].pad = d.boltz; ring[q].idx1 = 0x5eed; } // for thread q
                    }
                    if (w.lane == 0) { fl[0] = ok; fl[1] = cell; }
                }
                __syncthreads();
                accmask |= (unsigned long long)(fl[0] ? 1u : 0u) << q;
                if (fl[1]) cr = load_cell_regs(w); // the cell may have changed (also on a reverted shear / stretch)
                q += 1;
            }
            pairs_total += pairs;
            pairs = 0;
            if (my_type >= 0) { // thread t: move t of the batch
                const int my_acc = (int)((accmask >> threadIdx.x) & 1ull);
#pragma unroll
                for (int t = 0; t < 6; ++t) { att[t] += (my_type == t); acc[t] += (my_type == t && my_acc); }
                if (REPLAY && a.out) {
                    const PropRec *slot = ring + threadIdx.x;
                    int rs = my_acc ? HANS_ACCEPT : HANS_REJ_OVERLAP;
                    double bz = my_acc ? 1.0 : 0.0;
                    if (!my_plain && slot->idx1 == 0x5eed) { rs = slot->idx0; bz = slot->pad; }
                    hans_decision *o = a.out + (size_t)k * a.nmoves + m0 + threadIdx.x;
                    o->accepted = my_acc; o->reason = rs; o->boltz = bz;
                }
            }
            __syncthreads(); // the ring is rewritten by the next batch
        }
        {
            WCtx wl = w;
            wl.lane = threadIdx.x;
            store_walker<T>(wl, a.H, a.R, wi);
        }
        if (a.attempted) {
#pragma unroll
            for (int t = 0; t < 6; ++t) {
                unsigned at = att[t], ac = acc[t];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    at += __shfl_down_sync(FULL, at, o);
                    ac += __shfl_down_sync(FULL, ac, o);
                }
                if (w.lane == 0) {
                    atomicAdd(a.attempted + (size_t)k * 6 + t, (unsigned long long)at);
                    atomicAdd(a.accepted + (size_t)k * 6 + t, (unsigned long long)ac);
                }
            }
        }
        if (threadIdx.x == 0) {
            if (!REPLAY) a.ctr[wi] = ctr0 + (unsigned long long)a.nmoves;
            else if (a.ctr) a.ctr[wi] += (unsigned long long)a.nmoves;
            if (a.vol) a.vol[wi] = fabs(det3(smem + W_SH(w)));
        }
        if (a.pair_tests) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) pairs_total += __shfl_down_sync(FULL, pairs_total, o);
            if (w.lane == 0 && pairs_total) atomicAdd(a.pair_tests, pairs_total);
        }
        __syncthreads();
    }
}

} // namespace hb
