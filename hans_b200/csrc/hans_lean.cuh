// hans_lean.cuh -- MC_run, one warp per walker, with the beads-per-chain count a compile-time
// constant: the default walk kernel for walkers of up to 128 beads.
//
// What the profile of the generic cooperative kernel (hans_walk.cuh, profiles/r01_ncu_c2_*) said:
// ~1.7 warps per SM scheduler (the BASELINE configurations have ~1 000 walkers per GPU and a
// Markov chain is serial), ~400 warp instructions and ~2 300 cycles per single-sphere move, a
// 150 KB kernel whose hot part does not fit the 32 KB instruction cache ("no instruction" is the
// second largest stall).  The time of a launch is (instructions on a walker's critical path) x
// (latency per instruction), so this kernel
//   * keeps the cell inverse and the float32 filter parameters in registers between box moves,
//   * for single spheres lets every lane build the trial bead redundantly (no staging, no
//     synchronisation before the pair tests) and commits from registers,
//   * for chains keeps the trial chain's float32 shadows in registers (NB of them) and walks all
//     other beads in fully unrolled code, one vote (REDUX) per move,
//   * keeps everything that is not a single-chain move out of line, so the loop that runs >90% of
//     the moves is a few KB of straight-line code.
// Same arithmetic as hans_walk.cuh (make_trial, r2_f32, pair_r2, move_box are shared), hence the
// same decisions and trajectories, bit for bit; tests/test_gpu_parity.py runs both kernels against
// the CPU checker.
//
// Reference call sites replaced: NesSa/MCNS.py:304-383 (the move loop of MC_run).
#pragma once

#include "hans_walk.cuh"

namespace hb {

// Shared-memory accesses of the hottest path (single-sphere translation) through ONE 32-bit shared base kept in a
// register.  nvcc forms every address of the `extern __shared__` symbol with S2R SR_CgaCtaId + LEA; at the head of
// a move that chain is pure latency in front of the first load (profiles/r01_ncu_c2_lean.txt, SASS view).
__device__ __forceinline__ unsigned shared_base()
{
    unsigned b = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("" : "+r"(b)); // opaque: keep it in a register instead of recomputing it
    return b;
}
__device__ __forceinline__ double lds_f64(unsigned a)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ float4 lds_f4(unsigned a)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ int2 lds_i2(unsigned a)
{
    int2 v;
    asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f64(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;" :: "r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void sts_f4(unsigned a, float4 v)
{
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// cell inverse + filter parameters of the current cell, in registers
struct CellRegs {
    double Hi[9];
    Filt f;
};

__device__ __forceinline__ CellRegs load_cell_regs(const WCtx &w)
{
    CellRegs c;
#pragma unroll
    for (int k = 0; k < 9; ++k) c.Hi[k] = smem[W_SH(w) + 9 + k];
    c.f = load_filt(W_SF(w));
    return c;
}

// cold: the float64 settle of a trial chain that had a pair inside the filter's band
__device__ __noinline__ bool lean_recheck(WCtx w, int ichain, int b0, bool amb, bool amb_intra)
{
    return __any_sync(FULL, recheck_chain_exact<32>(w, ichain, b0, amb, amb_intra));
}

// cold: chain moves that are not on the fast path of their specialisation
__device__ __noinline__ DecP lean_move_chain_generic(WCtx w, int q)
{
    unsigned pr = 0;
    const Dec d = move_chain<32>(w, W_RING(w) + q, HANS_MODE_FUSED, pr);
    DecP o;
    o.accepted = d.accepted; o.reason = d.reason; o.boltz = d.boltz; o.pairs = pr;
    return o;
}

// single spheres (NB == 1), translation: alk.alkane_translate_chain + accept/restore
// (MCNS.py:323,371-380).  Returns 1 accepted / 0 rejected.  u_ok = the proposal's u_acc < 1 (MCNS.py:372 with boltz = 1:
// always true for generated proposals; decided when the record entered the ring).
__device__ __forceinline__ int lean_translate_sphere(const WCtx &w, const CellRegs &cr, unsigned sb, int q, int ichain,
                                                     bool u_ok, unsigned &pairs)
{
    const int N = w.N;
    const unsigned pa = sb + 64u * (unsigned)(w.ring + q);             // the proposal record
    const unsigned Pa = sb + 8u * (unsigned)(W_P(w) + 3 * ichain);     // the bead (float64)
    const unsigned fa = sb + 16u * (unsigned)W_F4(w);                  // the shadows
    const double x = lds_f64(Pa) + lds_f64(pa + 16), y = lds_f64(Pa + 8) + lds_f64(pa + 24), z = lds_f64(Pa + 16) + lds_f64(pa + 32);
    const float4 cs = shadow_of(x, y, z, cr.Hi);
    const Filt &f = cr.f;
    float rmin = 3.0e38f; // smallest float32 r^2 over the beads this lane owns
#pragma unroll 2
    for (int j = w.lane; j < N; j += 32) {
        const float4 s = lds_f4(fa + 16u * (unsigned)j);
        const float r2 = r2_f32(f, s.x - cs.x, s.y - cs.y, s.z - cs.z);
        rmin = (j != ichain) ? fminf(rmin, r2) : rmin;
    }
    pairs += (unsigned)((N - 1 - w.lane + 32) >> 5);
    // one REDUX for the vote: as integers all negative floats sort below all positive ones, and a (tiny,
    // rounding) negative r^2 is "< lo" like any other certain overlap
    const float m = __int_as_float(__reduce_min_sync(FULL, __float_as_int(rmin)));
    bool overlap = m < f.lo;
    if (!overlap && !(m > f.hi)) { // nothing overlaps for certain, but a pair is inside the band: float64 decides
        if (w.lane == 0) {
            const int cn = W_CN(w);
            smem[cn] = x; smem[cn + 1] = y; smem[cn + 2] = z;
        }
        __syncwarp();
        overlap = lean_recheck(w, ichain, 0, true, false);
    }
    const bool accept = !overlap && u_ok; // MCNS.py:372
    if (accept) {
        if (w.lane == 0) {
            sts_f64(Pa, x); sts_f64(Pa + 8, y); sts_f64(Pa + 16, z);
            sts_f4(fa + 16u * (unsigned)ichain, cs);
        }
        __syncwarp();
    }
    return accept ? 1 : 0;
}

// chains of NB beads: translate / rotate / dihedral; chain-level pruning, trial shadows in registers
template <int NB>
__device__ __forceinline__ int lean_move_chain(const WCtx &w, const CellRegs &cr, const PropRec *p, int ichain,
                                               bool u_ok, unsigned &pairs)
{
    const int c0 = ichain * NB, cn = W_CN(w), c4 = W_C4(w), f4 = W_F4(w);
    int b0 = 0;
    bool intra = false;
    make_trial(w, p, w.lane, 32, cn, c4, b0, intra, cr.Hi); // (proposal validated by the caller)
    __syncwarp();
    const float Rt = intra ? chain_radius_warp(cn, NB) : SMF[w.rad + ichain]; // translations and rotations are rigid
    const bool overlap = chain_overlaps_pruned<32, NB>(w, cr.f, ichain, b0, intra, Rt, pairs);
    const bool accept = !overlap && u_ok; // MCNS.py:372 with boltz = 1
    if (accept) {
        const int P = W_P(w) + 3 * c0;
        {   // beads (3 NB doubles = 3 NB / 2 sixteen-byte words; P and cn are even) and shadows in one pass of 16-byte copies
            constexpr int n2 = 3 * NB / 2;
            static_assert(NB % 2 == 0 && n2 + NB <= 32, "one pass");
            const int k = w.lane;
            if (k < n2 + NB) {
                const int src = k < n2 ? (cn >> 1) + k : c4 + (k - n2);
                const int dst = k < n2 ? (P >> 1) + k : f4 + c0 + (k - n2);
                SM4[dst] = SM4[src];
            }
        }
        if (intra && w.lane == 0) SMF[w.rad + ichain] = Rt;
    }
    __syncwarp();
    return accept ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------
// Speculative windows.  A Markov chain is serial, but a window of consecutive single-chain moves
// can be EVALUATED together: proposals do not depend on the state except through the chain they
// move, and a hard-body decision is "does the trial chain overlap any bead of another chain".
//   phase A  the trial chains of all moves of the window are built from the current state (one
//            lane per trial bead), then every lane tests the resident beads it owns (lane + 32 r)
//            against every trial bead of every move: bit (r W + k) of `ovw` = "my bead of pass r
//            certainly overlaps the trial chain of move k" (`ambw`: a pair inside the filter's band).
//            These are independent pair tests, several in flight per lane -- the instruction-level
//            parallelism a single serial move does not have;
//   phase B  ONE REDUX-OR over the warp gives the blocked moves of the whole window.  Moves are
//            settled in order: blocked moves before the first free one are rejected at no cost;
//            the free one is accepted and committed, which changes one chain c: the lanes owning
//            beads of c drop their bits for the later moves (those were against the old
//            position), the later moves re-test their trial chain against the new chain c only
//            (a few pairs per lane), and the window is cut before a later move of chain c itself
//            (its trial was built from the old position; it starts the next window).
// Every decision is thereby taken against exactly the state the serial order would have produced,
// so trajectories are bit-identical to the serial kernels and to the CPU checker.
// Returns the number of moves settled (>= 1); `accbits` = the accepted ones.
// ---------------------------------------------------------------------------------------------
template <int W> __device__ __forceinline__ unsigned fold_passes(unsigned x)
{
    x |= x >> (2 * W);
    x |= x >> W;
    return x & ((1u << W) - 1u);
}

template <int NB, int W>
__device__ __forceinline__ int lean_window(const WCtx &w, const CellRegs &cr, int q, int nw, unsigned &accbits,
                                           unsigned &pairs)
{
    static_assert(W == 4 || W == 8, "four passes of W bits share one 32-bit word");
    const PropRec *ring = W_RING(w) + q;
    const int lane = w.lane, N = w.N, f4 = W_F4(w), np = (N + 31) >> 5;
    const Filt &f = cr.f;

    // ---- phase A ------------------------------------------------------------------------------
    for (int e = lane; e < nw * NB; e += 32) { // one lane per trial bead
        const int k = e / NB, b = e - k * NB;
        int b0;
        bool intra;
        make_trial<false>(w, ring + k, b, NB, w.wc + 3 * NB * k, w.w4 + NB * k, b0, intra);
    }
    const int myck = lane < nw ? ring[lane].ichain : -1; // lane k remembers the chain of move k
    __syncwarp();
    unsigned ovw = 0, ambw = 0;   // bit r W + k, against the state at the start of the window
    unsigned fixw = 0, afixw = 0; // bit k: intra-chain pairs, and pairs against chains committed in this window
    unsigned dih = 0;
    for (int r0 = 0; r0 < np; r0 += 2) {
        const int j0 = lane + 32 * r0, j1 = j0 + 32;
        const bool live0 = j0 < N, live1 = j1 < N;
        const float4 s0 = SM4[f4 + (live0 ? j0 : 0)], s1 = SM4[f4 + (live1 ? j1 : 0)];
        const int cj0 = j0 / NB, cj1 = j1 / NB;
#pragma unroll 2
        for (int k = 0; k < nw; ++k) {
            const int4 hd = *reinterpret_cast<const int4 *>(ring + k);
            const int b0 = hd.x == HANS_DIH ? hd.z : 0;
            if (hd.x == HANS_DIH) dih |= 1u << k;
            bool o0 = false, o1 = false, a0 = false, a1 = false;
#pragma unroll(NB <= 2 ? NB : 1)
            for (int b = b0; b < NB; ++b) {
                const float4 t = SM4[w.w4 + NB * k + b];
                const float ra = r2_f32(f, s0.x - t.x, s0.y - t.y, s0.z - t.z);
                const float rb = r2_f32(f, s1.x - t.x, s1.y - t.y, s1.z - t.z);
                o0 |= ra < f.lo; a0 |= !(ra > f.hi);
                o1 |= rb < f.lo; a1 |= !(rb > f.hi);
            }
            const bool on0 = live0 && cj0 != hd.y, on1 = live1 && cj1 != hd.y;
            ovw |= ((on0 && o0) ? 1u : 0u) << (r0 * W + k) | ((on1 && o1) ? 1u : 0u) << (r0 * W + W + k);
            ambw |= ((on0 && a0) ? 1u : 0u) << (r0 * W + k) | ((on1 && a1) ? 1u : 0u) << (r0 * W + W + k);
            pairs += (unsigned)(((on0 ? 1 : 0) + (on1 ? 1 : 0)) * (NB - b0));
        }
    }
    if (NB >= 4 && dih) { // dihedral moves: intra-chain pairs (j, b), b >= b0, beyond the exclusion
        const uchar2 *tab = W_ITAB(w);
        for (int k = 0; k < nw; ++k) {
            if (!((dih >> k) & 1u)) continue;
            const int b0 = ring[k].idx0, t4 = w.w4 + NB * k;
            for (int e = lane; e < w.npc; e += 32) {
                const uchar2 jb = tab[e];
                if ((int)jb.y >= b0) {
                    const float4 a = SM4[t4 + jb.x], b = SM4[t4 + jb.y];
                    const float r2 = r2_f32(f, b.x - a.x, b.y - a.y, b.z - a.z);
                    if (r2 < f.lo) fixw |= 1u << k;
                    if (!(r2 > f.hi)) afixw |= 1u << k;
                    pairs += 1;
                }
            }
        }
    }

    // ---- phase B ------------------------------------------------------------------------------
    accbits = 0;
    unsigned pend = (nw >= 32) ? FULL : ((1u << nw) - 1u);
    constexpr int LB = 32 / W; // lanes per later move in the fix-ups
    for (;;) {
        const unsigned blocked = fold_passes<W>(__reduce_or_sync(FULL, ovw | fixw));
        const unsigned cand = pend & ~blocked;
        if (!cand) break; // every move still pending overlaps something: rejected
        const unsigned band = fold_passes<W>(__reduce_or_sync(FULL, ambw | afixw));
        const int i = __ffs(cand) - 1;
        pend &= ~((2u << i) - 1u);
        const PropRec *pi = ring + i;
        const int ci = pi->ichain, b0i = pi->type == HANS_DIH ? pi->idx0 : 0;
        const int wci = w.wc + 3 * NB * i, w4i = w.w4 + NB * i;
        if ((band >> i) & 1u) { // no certain overlap, but a pair inside the band: float64 decides
            const int cn = W_CN(w);
            for (int e = lane; e < 3 * NB; e += 32) smem[cn + e] = smem[wci + e];
            __syncwarp();
            if (lean_recheck(w, ci, b0i, true, pi->type == HANS_DIH)) continue;
        }
        if (!(pi->u_acc < 1.0)) continue; // MCNS.py:372 with boltz = 1 (only explicit proposals can fail it): rejected
        accbits |= 1u << i;
        {
            const int P = W_P(w) + 3 * NB * ci, f4c = f4 + NB * ci;
            for (int e = lane; e < 3 * NB; e += 32) smem[P + e] = smem[wci + e];
            for (int e = lane; e < NB; e += 32) SM4[f4c + e] = SM4[w4i + e];
            if (NB > 1 && pi->type == HANS_DIH && lane == 0) SMF[w.rad + ci] = chain_radius(wci, NB);
        }
        __syncwarp();
        if (!pend) break;
        // a later move of the same chain was built from the old position: it starts the next window
        const unsigned stale = __ballot_sync(FULL, lane > i && myck == ci) & pend;
        if (stale) {
            nw = __ffs(stale) - 1;
            pend &= (1u << nw) - 1u;
            if (!pend) break;
        }
        // my beads of chain ci were tested at their old position: those bits are void
#pragma unroll
        for (int r = 0; r < 4; ++r)
            if ((lane + 32 * r) / NB == ci) { ovw &= ~(((1u << W) - 1u) << (r * W)); ambw &= ~(((1u << W) - 1u) << (r * W)); }
        // the later moves against the new chain ci
        {
            const int k = i + 1 + lane / LB, sub = lane % LB;
            if (k < nw) {
                const int b0 = ring[k].type == HANS_DIH ? ring[k].idx0 : 0;
                bool o = false, am = false;
                for (int bb = sub; bb < NB; bb += LB) {
                    const float4 sn = SM4[f4 + NB * ci + bb];
                    for (int b = b0; b < NB; ++b) {
                        const float4 t = SM4[w.w4 + NB * k + b];
                        const float r2 = r2_f32(f, sn.x - t.x, sn.y - t.y, sn.z - t.z);
                        o |= r2 < f.lo; am |= !(r2 > f.hi);
                    }
                    pairs += (unsigned)(NB - b0);
                }
                if (o) fixw |= 1u << k;
                if (am) afixw |= 1u << k;
            }
        }
    }
    return nw;
}

// W = 1: strictly one move at a time; W = 8: speculative windows of up to W moves.
// MINB = resident one-warp CTAs per SM the register allocation must allow.  1: no cap (214 - 221 registers: 9 walkers per
// SM), fastest per walker; 12: 168 registers; 16: 128 registers -- slower per walker (C3 -15 % / -23 %, C2 -11 % / -22 % at
// 1 024 walkers per GPU), more walkers in flight.  The launcher picks by the size of the batch (launch_lean).
template <int NB, int W, bool REPLAY, int MINB>
__global__ void __launch_bounds__(32, MINB) walk_lean_kernel(const WalkArgs a)
{
    constexpr int T = 32;
    const WCtx w = setup_ctx_g<32, 1>(a.cfg, W > 1 ? W : 0);
    const double limit = a.d_limit ? *a.d_limit : a.limit;
    for (int k = blockIdx.x; k < a.n_active; k += gridDim.x) {
        const int wi = a.ids ? a.ids[k] : k;
        load_walker<T>(w, a.H, a.R, wi, true);
        CellRegs cr = load_cell_regs(w);
        const unsigned sb = shared_base();
        const unsigned long long ctr0 = REPLAY ? 0ull : a.ctr[wi];
        unsigned att[6] = {0, 0, 0, 0, 0, 0}, acc[6] = {0, 0, 0, 0, 0, 0};
        unsigned pairs = 0;
        unsigned long long pairs_total = 0;
        PropRec *ring = W_RING(w);
        for (int m0 = 0; m0 < a.nmoves; m0 += T) {
            const int nbatch = a.nmoves - m0 < T ? a.nmoves - m0 : T;
            int my_type = -1;
            bool my_plain = false; // a single-chain move with a well-formed proposal
            bool my_uok = false;   // u_acc < 1: an accepted single-chain move commits (MCNS.py:372 with boltz = 1)
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
                prep_slot(slot);
                my_uok = slot->u_acc < 1.0;
                my_plain = (my_type == HANS_TRANS || (NB > 1 && my_type == HANS_ROT) ||
                            (my_type == HANS_DIH && NB >= 4 && slot->idx0 >= 3 && slot->idx0 <= NB - 1)) &&
                           (unsigned)slot->ichain < (unsigned)w.nc;
            }
            __syncwarp();
            const unsigned plain = W > 1 ? __ballot_sync(FULL, my_plain) : 0u;
            const unsigned uok = __ballot_sync(FULL, my_uok);
            // decisions of the batch: accepted bits (uniform); reason / boltz of the moves that are not
            // plain single-chain moves are kept by lane q (a chain move's follow from its bit)
            unsigned accmask = 0;
            int my_reason = -1;
            double my_boltz = 0.0;
            for (int q = 0; q < nbatch; ++q) {
                if (W > 1) { // a run of at least two plain single-chain moves: settle it as a window
                    const unsigned run = ~(plain >> q);
                    int nw = run ? __ffs(run) - 1 : 32;
                    if (nw > W) nw = W;
                    if (nw >= 2) {
                        unsigned ab;
                        const int done = lean_window<NB, (W > 1 ? W : 4)>(w, cr, q, nw, ab, pairs);
                        accmask |= ab << q;
                        q += done - 1;
                        continue;
                    }
                }
                const PropRec *p = ring + q;
                int type, ichain, idx0;
                if (NB == 1) {
                    const int2 h2 = lds_i2(sb + 64u * (unsigned)(w.ring + q)); // type, ichain
                    type = h2.x; ichain = h2.y; idx0 = 0;
                } else {
                    const int4 hd = *reinterpret_cast<const int4 *>(p); // type, ichain, idx0, idx1
                    type = hd.x; ichain = hd.y; idx0 = hd.z;
                }
                const bool chain_ok = (unsigned)ichain < (unsigned)w.nc;
                if (NB == 1 && type == HANS_TRANS && chain_ok) {
                    accmask |= (unsigned)lean_translate_sphere(w, cr, sb, q, ichain, ((uok >> q) & 1u) != 0u, pairs) << q;
                } else if (NB > 1 && chain_ok &&
                           (type == HANS_TRANS || type == HANS_ROT || (type == HANS_DIH && NB >= 4 && idx0 >= 3 && idx0 <= NB - 1))) {
                    accmask |= (unsigned)lean_move_chain<(NB > 1 ? NB : 2)>(w, cr, p, ichain, ((uok >> q) & 1u) != 0u, pairs) << q;
                } else {
                    DecP d;
                    if (type >= HANS_TRANS && type <= HANS_DIH) {
                        d = lean_move_chain_generic(w, q);
                    } else {
                        d = move_box<T>(w, q, limit, HANS_MODE_FUSED);
                        cr = load_cell_regs(w); // the cell may have changed (also on a reverted shear / stretch)
                    }
                    pairs += d.pairs;
                    accmask |= (d.accepted ? 1u : 0u) << q;
                    if (w.lane == q) { my_reason = d.reason; my_boltz = d.boltz; }
                }
            }
            pairs_total += pairs;
            pairs = 0;
            const int my_acc = (int)((accmask >> w.lane) & 1u);
            if (my_reason < 0) { my_reason = my_acc ? HANS_ACCEPT : HANS_REJ_OVERLAP; my_boltz = my_acc ? 1.0 : 0.0; }
            if (my_type >= 0) {
#pragma unroll
                for (int t = 0; t < 6; ++t) { att[t] += (my_type == t); acc[t] += (my_type == t && my_acc); }
                if (REPLAY && a.out) {
                    hans_decision *o = a.out + (size_t)k * a.nmoves + m0 + w.lane;
                    o->accepted = my_acc; o->reason = my_reason; o->boltz = my_boltz;
                }
            }
            __syncwarp(); // the ring is rewritten by the next batch
        }
        store_walker<T>(w, a.H, a.R, wi);
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
        if (w.lane == 0) {
            if (!REPLAY) a.ctr[wi] = ctr0 + (unsigned long long)a.nmoves;
            else if (a.ctr) a.ctr[wi] += a.ctr_add; // proposals came from hans_gen_proposals
            if (a.vol) a.vol[wi] = fabs(det3(smem + W_SH(w)));
        }
        if (a.pair_tests) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) pairs_total += __shfl_down_sync(FULL, pairs_total, o);
            if (w.lane == 0 && pairs_total) atomicAdd(a.pair_tests, pairs_total);
        }
        __syncwarp();
    }
}

} // namespace hb
