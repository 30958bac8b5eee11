// hans_spec.cuh -- MC_run with SPECULATIVE WINDOWS: one warp per walker, several moves in flight.
//
// Why: the configurations BASELINE.json names have ~1 000 walkers per GPU, i.e. ~1.7 warps per SM
// scheduler when a warp owns a walker, and a Markov chain is serial: the cooperative kernel
// (hans_walk.cuh) spends ~400 warp instructions and ~2 300 cycles per single-sphere move on
// 2 pair tests per lane (ncu, profiles/r01_*).  Parallelism has to come from INSIDE the chain.
//
// How: single-chain proposals (translate / rotate / dihedral) do not depend on the state except
// through the chain they move, and a hard-body decision is "does the trial chain overlap any other
// chain".  For a window of W = 32/L consecutive single-chain moves, L lanes per move:
//   phase A  every move builds its trial chain from the CURRENT state and computes, against all
//            other chains, a bit mask of the chains it overlaps (plus its own bit for the
//            intra-chain pairs of a dihedral move) -- all W moves at once, no synchronisation;
//   phase B  the moves are settled in order.  A move with an empty mask is accepted; committing it
//            changes exactly one chain c, so every later move of the window re-tests its trial
//            chain against the new chain c only (one bit of its mask), and later moves of chain c
//            itself are marked stale and re-evaluated cooperatively when their turn comes.
// Masks are therefore exact with respect to the state each move would have seen in the serial
// order, so decisions and trajectories are BIT-IDENTICAL to the serial kernel and to the CPU
// checker (tests/test_gpu_parity.py runs both against it).  Box moves (volume / shear / stretch)
// end a window and run on the cooperative code of hans_walk.cuh.
//
// Reference call sites replaced: NesSa/MCNS.py:304-383 (the move loop of MC_run).
#pragma once

#include "hans_walk.cuh"

namespace hb {

// float32 filter + float64 settle of ONE pair: trial bead (shadow c, Cartesian at smem[cb..]) against
// resident bead j (shadow SM4[f4j], Cartesian at smem[pj..])
__device__ __forceinline__ bool pair_hits(const Filt &f, const float4 c, int cb, int f4j, int pj, int sh)
{
    const float4 s = SM4[f4j];
    const float r2 = r2_f32(f, s.x - c.x, s.y - c.y, s.z - c.z);
    bool hit = r2 < f.lo;
    if (!hit && !(r2 > f.hi)) { // inside the band (~1e-4 of the pairs): the float64 arithmetic decides
        hit = pair_r2(smem + sh, smem + sh + 9, smem[pj] - smem[cb], smem[pj + 1] - smem[cb + 1],
                      smem[pj + 2] - smem[cb + 2]) < 1.0;
        atomicAdd(&g_diag[2], 1ull);
    }
    return hit;
}

// trial beads [b0, nb) of a move against the beads k = kfirst, kfirst + kstride, ... of chain c
__device__ __forceinline__ bool trial_hits_chain(const WCtx &w, const Filt &f, int wc, int w4, int b0, int c,
                                                 int kfirst, int kstride, unsigned &pairs)
{
    const int nb = w.nb, f4 = W_F4(w) + c * nb, p = W_P(w) + 3 * c * nb, sh = W_SH(w);
    bool hit = false;
    for (int b = b0; b < nb; ++b) {
        const float4 cs = SM4[w4 + b];
        for (int k = kfirst; k < nb; k += kstride) hit |= pair_hits(f, cs, wc + 3 * b, f4 + k, p + 3 * k, sh);
    }
    pairs += (unsigned)((nb - b0) * ((nb - kfirst + kstride - 1) / kstride));
    return hit;
}

// One window: ring entries [q, q + nw) are single-chain moves with valid proposals, 2 <= nw <= 32/L.
// Returns the accepted moves as a bit mask over the window.
template <int L>
__device__ __forceinline__ unsigned spec_window(const WCtx &w, int q, int nw, unsigned &pairs)
{
    const int lane = w.lane, mv = lane / L, sub = lane % L;
    const bool on = mv < nw;
    const PropRec *ring = W_RING(w);
    const PropRec *p = ring + q + (on ? mv : 0);
    const int nb = w.nb, nc = w.nc, ichain = p->ichain;
    const int wc = w.wc + 3 * nb * mv, w4 = w.w4 + nb * mv, sh = W_SH(w);
    const Filt f = load_filt(W_SF(w));
    int b0 = 0;
    bool intra = false;

    // ---- phase A: trial chains and overlap masks of all moves of the window -------------------
    if (on) make_trial(w, p, sub, L, wc, w4, b0, intra);
    __syncwarp();
    unsigned mlo = 0, mhi = 0; // chains 0..31 / 32..63 this trial chain overlaps
    if (on) {
        if (nb == 1) { // single spheres: one pair per chain, straight-line code
            const float4 cs = SM4[w4];
            const int f4 = W_F4(w), P = W_P(w);
            unsigned bit = 1u << sub;
#pragma unroll 4
            for (int c = sub; c < nc && c < 32; c += L, bit <<= L)
                if (pair_hits(f, cs, wc, f4 + c, P + 3 * c, sh) && c != ichain) mlo |= bit;
            bit = 1u << sub;
#pragma unroll 4
            for (int c = 32 + sub; c < nc; c += L, bit <<= L)
                if (pair_hits(f, cs, wc, f4 + c, P + 3 * c, sh) && c != ichain) mhi |= bit;
            pairs += (unsigned)((nc - sub + L - 1) / L);
        } else {
            for (int c = sub; c < nc; c += L) {
                if (c == ichain) continue;
                if (trial_hits_chain(w, f, wc, w4, b0, c, 0, 1, pairs)) {
                    if (c < 32) mlo |= 1u << c; else mhi |= 1u << (c - 32);
                }
            }
            if (intra) { // dihedral: pairs (j, b), b >= b0, beyond the exclusion -> the move's own bit
                const uchar2 *tab = W_ITAB(w);
                bool hit = false;
                for (int e = sub; e < w.npc; e += L) {
                    const uchar2 jb = tab[e];
                    if ((int)jb.y >= b0) {
                        hit |= pair_hits(f, SM4[w4 + jb.y], wc + 3 * jb.y, w4 + jb.x, wc + 3 * jb.x, sh);
                        pairs += 1;
                    }
                }
                if (hit) { if (ichain < 32) mlo |= 1u << ichain; else mhi |= 1u << (ichain - 32); }
            }
        }
    }
#pragma unroll
    for (int o = 1; o < L; o <<= 1) {
        mlo |= __shfl_xor_sync(FULL, mlo, o);
        mhi |= __shfl_xor_sync(FULL, mhi, o);
    }

    // ---- phase B: settle in order; only accepted and stale moves cost an iteration ------------
    unsigned accbits = 0;
    bool stale = false;
    int cur = 0;
    const bool leader = on && sub == 0;
    for (;;) {
        const bool pending = leader && mv >= cur;
        const unsigned cand = __ballot_sync(FULL, pending && (stale || (mlo | mhi) == 0u));
        if (!cand) break;
        const unsigned stl = __ballot_sync(FULL, pending && stale);
        const int src = __ffs(cand) - 1, i = src / L;
        const PropRec *pi = ring + q + i;
        const int ci = pi->ichain;
        bool ok = true, commit;
        if ((stl >> src) & 1u) { // chain ci was moved earlier in this window: evaluate from scratch
            const Dec d = move_chain<32>(w, pi, HANS_MODE_FUSED, pairs);
            ok = d.accepted != 0;
            commit = ok && (pi->u_acc < 1.0);
        } else {
            commit = pi->u_acc < 1.0; // MCNS.py:372 with boltz = 1
            if (commit) {
                const int P = W_P(w) + 3 * ci * nb, f4 = W_F4(w) + ci * nb;
                const int wci = w.wc + 3 * nb * i, w4i = w.w4 + nb * i;
                for (int k = lane; k < 3 * nb; k += 32) smem[P + k] = smem[wci + k];
                for (int k = lane; k < nb; k += 32) SM4[f4 + k] = SM4[w4i + k];
            }
        }
        accbits |= (ok ? 1u : 0u) << i;
        cur = i + 1;
        if (commit) {
            __syncwarp();
            const bool later = on && mv >= cur && !stale;
            if (later && ichain == ci) stale = true;
            const bool fix = later && !stale;
            bool hit = false;
            if (fix) hit = trial_hits_chain(w, f, wc, w4, b0, ci, sub, L, pairs);
#pragma unroll
            for (int o = 1; o < L; o <<= 1) hit |= __shfl_xor_sync(FULL, hit ? 1 : 0, o) != 0;
            if (fix) {
                if (ci < 32) mlo = (mlo & ~(1u << ci)) | ((hit ? 1u : 0u) << ci);
                else mhi = (mhi & ~(1u << (ci - 32))) | ((hit ? 1u : 0u) << (ci - 32));
            }
        }
    }
    __syncwarp();
    return accbits;
}

// MC_run (MCNS.py:276-392) for a batch of walkers, one warp (= one CTA) per walker
template <int L, bool REPLAY>
__global__ void __launch_bounds__(32) walk_spec_kernel(const WalkArgs a)
{
    constexpr int T = 32, W = 32 / L;
    const WCtx w = setup_ctx_g<32, 1>(a.cfg, W);
    const double limit = a.d_limit ? *a.d_limit : a.limit;
    const int nb = w.nb;
    for (int k = blockIdx.x; k < a.n_active; k += gridDim.x) {
        const int wi = a.ids ? a.ids[k] : k;
        load_walker<T>(w, a.H, a.R, wi, true);
        const unsigned long long ctr0 = REPLAY ? 0ull : a.ctr[wi];
        unsigned att[6] = {0, 0, 0, 0, 0, 0}, acc[6] = {0, 0, 0, 0, 0, 0};
        unsigned pairs = 0;
        unsigned long long pairs_total = 0;
        PropRec *ring = W_RING(w);
        for (int m0 = 0; m0 < a.nmoves; m0 += T) {
            const int nbatch = a.nmoves - m0 < T ? a.nmoves - m0 : T;
            int my_type = -1;
            bool my_plain = false; // a single-chain move with a well-formed proposal
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
                my_plain = my_type == HANS_TRANS || my_type == HANS_ROT ||
                           (my_type == HANS_DIH && nb >= 4 && slot->idx0 >= 3 && slot->idx0 <= nb - 1);
                my_plain = my_plain && (unsigned)slot->ichain < (unsigned)w.nc;
            }
            __syncwarp();
            const unsigned plain = __ballot_sync(FULL, my_plain);
            int my_acc = 0, my_reason = 0;
            double my_boltz = 0.0;
            int q = 0;
            while (q < nbatch) {
                const unsigned run = ~(plain >> q);       // first zero of `plain` from q on
                int nw = run ? __ffs(run) - 1 : 32;       // (plain >> q has zeros above bit 31 - q)
                if (nw > W) nw = W;
                if (nw >= 2) {
                    const unsigned ab = spec_window<L>(w, q, nw, pairs);
                    const int rel = w.lane - q;
                    if (rel >= 0 && rel < nw) {
                        my_acc = (int)((ab >> rel) & 1u);
                        my_reason = my_acc ? HANS_ACCEPT : HANS_REJ_OVERLAP;
                        my_boltz = my_acc ? 1.0 : 0.0;
                    }
                    q += nw;
                    continue;
                }
                const int type = ring[q].type;
                int accepted, reason;
                double boltz;
                if (type >= HANS_TRANS && type <= HANS_DIH) {
                    const Dec d = move_chain<T>(w, ring + q, HANS_MODE_FUSED, pairs);
                    accepted = d.accepted; reason = d.reason; boltz = d.boltz;
                } else {
                    const DecP d = move_box<T>(w, q, limit, HANS_MODE_FUSED);
                    accepted = d.accepted; reason = d.reason; boltz = d.boltz; pairs += d.pairs;
                }
                if (w.lane == q) { my_acc = accepted; my_reason = reason; my_boltz = boltz; }
                ++q;
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
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    att[t] += __shfl_down_sync(FULL, att[t], o);
                    acc[t] += __shfl_down_sync(FULL, acc[t], o);
                }
                if (w.lane == 0) {
                    atomicAdd(a.attempted + (size_t)k * 6 + t, (unsigned long long)att[t]);
                    atomicAdd(a.accepted + (size_t)k * 6 + t, (unsigned long long)acc[t]);
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
