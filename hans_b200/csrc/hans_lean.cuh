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
__device__ __noinline__ Dec lean_move_chain_generic(WCtx w, int q, unsigned *pairs)
{
    unsigned pr = 0;
    const Dec d = move_chain<32>(w, W_RING(w) + q, HANS_MODE_FUSED, pr);
    *pairs += pr;
    return d;
}

// single spheres (NB == 1), translation: alk.alkane_translate_chain + accept/restore
// (MCNS.py:323,371-380).  Returns 1 accepted / 0 overlap.
__device__ __forceinline__ int lean_translate_sphere(const WCtx &w, const CellRegs &cr, const PropRec *p, int ichain,
                                                     unsigned &pairs)
{
    const int P = W_P(w) + 3 * ichain, f4 = W_F4(w), N = w.N;
    const double x = smem[P] + p->p[0], y = smem[P + 1] + p->p[1], z = smem[P + 2] + p->p[2];
    const float4 cs = shadow_of(x, y, z, cr.Hi);
    const Filt &f = cr.f;
    bool ov = false, amb = false;
#pragma unroll 2
    for (int j = w.lane; j < N; j += 32) {
        const float4 s = SM4[f4 + j];
        const float r2 = r2_f32(f, s.x - cs.x, s.y - cs.y, s.z - cs.z);
        const bool other = j != ichain;
        ov |= other && (r2 < f.lo);
        amb |= other && !(r2 > f.hi);
    }
    pairs += (unsigned)((N - 1 - w.lane + 32) >> 5);
    const unsigned code = __reduce_or_sync(FULL, (ov ? 2u : 0u) | (amb ? 1u : 0u));
    bool overlap = (code & 2u) != 0u;
    if (code == 1u) { // nothing overlaps for certain, but a pair is inside the band: float64 decides
        if (w.lane == 0) {
            const int cn = W_CN(w);
            smem[cn] = x; smem[cn + 1] = y; smem[cn + 2] = z;
        }
        __syncwarp();
        overlap = lean_recheck(w, ichain, 0, amb, false);
    }
    if (!overlap && (p->u_acc < 1.0)) { // MCNS.py:372 with boltz = 1
        if (w.lane == 0) {
            smem[P] = x; smem[P + 1] = y; smem[P + 2] = z;
            SM4[f4 + ichain] = cs;
        }
        __syncwarp();
    }
    return overlap ? 0 : 1;
}

// chains of NB beads: translate / rotate / dihedral with the trial shadows in registers
template <int NB>
__device__ __forceinline__ int lean_move_chain(const WCtx &w, const CellRegs &cr, const PropRec *p, int ichain,
                                               unsigned &pairs)
{
    const int c0 = ichain * NB, cn = W_CN(w), c4 = W_C4(w), f4 = W_F4(w), N = w.N;
    int b0 = 0;
    bool intra = false;
    make_trial(w, p, w.lane, 32, cn, c4, b0, intra); // (proposal validated by the caller)
    __syncwarp();
    float4 c[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) c[b] = SM4[c4 + b];
    const Filt &f = cr.f;
    bool ov = false, amb = false;
    for (int j = w.lane; j < N; j += 32) {
        if ((unsigned)(j - c0) >= (unsigned)NB) {
            const float4 s = SM4[f4 + j];
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                const float r2 = r2_f32(f, s.x - c[b].x, s.y - c[b].y, s.z - c[b].z);
                const bool moved = b >= b0;
                ov |= moved && (r2 < f.lo);
                amb |= moved && !(r2 > f.hi);
            }
            pairs += (unsigned)(NB - b0);
        }
    }
    bool amb_intra = false;
    if (intra) {
        const uchar2 *tab = W_ITAB(w);
        for (int e = w.lane; e < w.npc; e += 32) {
            const uchar2 jb = tab[e];
            if ((int)jb.y >= b0) {
                const float4 a = SM4[c4 + jb.x], b = SM4[c4 + jb.y];
                const float r2 = r2_f32(f, b.x - a.x, b.y - a.y, b.z - a.z);
                ov |= (r2 < f.lo);
                amb_intra |= !(r2 > f.hi);
                pairs += 1;
            }
        }
    }
    const unsigned code = __reduce_or_sync(FULL, (ov ? 4u : 0u) | (amb ? 1u : 0u) | (amb_intra ? 2u : 0u));
    bool overlap = (code & 4u) != 0u;
    if (!overlap && code) overlap = lean_recheck(w, ichain, b0, amb, amb_intra);
    if (!overlap && (p->u_acc < 1.0)) {
        const int P = W_P(w) + 3 * c0;
        for (int k = w.lane; k < 3 * NB; k += 32) smem[P + k] = smem[cn + k];
        for (int k = w.lane; k < NB; k += 32) SM4[f4 + c0 + k] = SM4[c4 + k];
    }
    __syncwarp();
    return overlap ? 0 : 1;
}

template <int NB, bool REPLAY>
__global__ void __launch_bounds__(32) walk_lean_kernel(const WalkArgs a)
{
    constexpr int T = 32;
    const WCtx w = setup_ctx_g<32, 1>(a.cfg, 0);
    const double limit = a.d_limit ? *a.d_limit : a.limit;
    for (int k = blockIdx.x; k < a.n_active; k += gridDim.x) {
        const int wi = a.ids ? a.ids[k] : k;
        load_walker<T>(w, a.H, a.R, wi, true);
        CellRegs cr = load_cell_regs(w);
        const unsigned long long ctr0 = REPLAY ? 0ull : a.ctr[wi];
        unsigned att[6] = {0, 0, 0, 0, 0, 0}, acc[6] = {0, 0, 0, 0, 0, 0};
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
            __syncwarp();
            int my_acc = 0, my_reason = 0;
            double my_boltz = 0.0;
            for (int q = 0; q < nbatch; ++q) {
                const PropRec *p = ring + q;
                const int4 hd = *reinterpret_cast<const int4 *>(p); // type, ichain, idx0, idx1
                const int type = hd.x, ichain = hd.y;
                int accepted, reason;
                double boltz;
                const bool chain_ok = (unsigned)ichain < (unsigned)w.nc;
                if (NB == 1 && type == HANS_TRANS && chain_ok) {
                    accepted = lean_translate_sphere(w, cr, p, ichain, pairs);
                    reason = accepted ? HANS_ACCEPT : HANS_REJ_OVERLAP; boltz = accepted ? 1.0 : 0.0;
                } else if (NB > 1 && chain_ok &&
                           (type == HANS_TRANS || type == HANS_ROT || (type == HANS_DIH && NB >= 4 && hd.z >= 3 && hd.z <= NB - 1))) {
                    accepted = lean_move_chain<(NB > 1 ? NB : 2)>(w, cr, p, ichain, pairs);
                    reason = accepted ? HANS_ACCEPT : HANS_REJ_OVERLAP; boltz = accepted ? 1.0 : 0.0;
                } else if (type >= HANS_TRANS && type <= HANS_DIH) {
                    const Dec d = lean_move_chain_generic(w, q, &pairs);
                    accepted = d.accepted; reason = d.reason; boltz = d.boltz;
                } else {
                    const DecP d = move_box<T>(w, q, limit, HANS_MODE_FUSED);
                    accepted = d.accepted; reason = d.reason; boltz = d.boltz; pairs += d.pairs;
                    cr = load_cell_regs(w); // the cell may have changed (also on a reverted shear / stretch)
                }
                if (w.lane == q) { my_acc = accepted; my_reason = reason; my_boltz = boltz; }
            }
            pairs_total += pairs;
            pairs = 0;
            if (my_type >= 0) {
#pragma unroll
                for (int t = 0; t < 6; ++t) { att[t] += (my_type == t); acc[t] += (my_type == t && my_acc); }
                if (REPLAY) {
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
