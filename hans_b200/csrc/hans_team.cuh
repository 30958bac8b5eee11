// hans_team.cuh -- MC_run for LARGE walkers (hundreds of beads): one CTA of NWARP warps per walker,
// up to NWARP consecutive single-chain moves in flight, ONE WARP PER MOVE.
//
// Why: the BASELINE shapes with large walkers have few of them per GPU (C4 512, C5 256), so the
// chip is short of independent instruction streams, and the cooperative kernel (hans_walk.cuh: all
// threads of the CTA share every move) spends a move on four barriers with one wave of work between
// them (profiles/r01_ncu_c5_cooperative_pruned.txt: 427 instructions per warp and move, barrier
// stall 2.1 of 7.8 cycles per instruction).  A Markov chain is serial, but a hard-body decision is
// "does the trial chain overlap a bead of another chain", and with 32..128 chains per walker the
// chains of consecutive moves are almost always distinct.  So a window of up to NWARP consecutive
// plain chain moves of DISTINCT chains is evaluated at once:
//   A1  warp k builds the trial chain of move k from the current state (its chain is not touched
//       by moves 0..k-1 of the window, so this is the state the serial order would have used);
//   A2  warp k tests its trial chain, with chain-level pruning and the float32 filter, against
//       every chain that no earlier move of the window touches (these are where the serial order
//       would have them);
//   --  barrier
//   A3  ... and against the chains of the earlier moves j < k twice: at their old position (bit j
//       of `old`) and at their trial position (bit j of `new`), bounding spheres first;
//   --  barrier
//   B   every thread settles the window in order (a few integer operations): move k overlaps iff
//       its base test did, or for some j < k: committed[j] ? new[j] : old[j].  A move with a pair
//       inside the filter's band cuts the window and is re-done by the cooperative float64 path;
//   C   the warps of committed moves copy their trial chains into the state; barrier.
// Three barriers per window instead of four per move, NWARP independent instruction streams per
// walker, no redundant work except the 2 nb^2 pair tests per (move, earlier move).  Every decision
// is taken against exactly the state the serial order would have produced, so the trajectory is
// bit-identical to the other kernels and to the CPU checker (tests/test_gpu_parity.py).
// Box moves and malformed proposals go through the cooperative routines (move_box<T>, move_chain<T>).
//
// Reference call sites replaced: NesSa/MCNS.py:304-383 (the move loop of MC_run).
#pragma once

#include "hans_walk.cuh"

namespace hb {

// behind the window area (int offsets from `res`): res[8] result word per move of the window; rtw[8] (at 24) bounding
// radii of the window's trial chains; meta[T + 8] (at 32) per move of the batch: chain | (u_acc < 1) << 30 for a plain
// single-chain move, -1 otherwise; info[T] (at 40 + T) per move: length of the window that starts there | commit mask
// << 8; then (at 48 + 2 T) one near list of TEAM_LIST chains per warp; then 32 float4 per warp: the shadows of the
// beads of one pass that lie inside the trial chain's bounding sphere
#define TEAM_WAVES 4
#define TEAM_LIST (32 * TEAM_WAVES)
#ifndef HANS_TEAM_BEAD_PRUNE
#define HANS_TEAM_BEAD_PRUNE 1
#endif
__host__ __device__ constexpr int team_tail_bytes(int nwarp) { return 192 + 256 * nwarp + 4 * TEAM_LIST * nwarp + 512 * nwarp; }

// A2 for move k of the window (warp-uniform arguments; called by ALL warps, `active` = this warp has a move: there
// is a CTA barrier inside).  `wmeta` = the window's entries of meta[].  Result word, uniform over the warp:
// bit 0 certain overlap with a chain no earlier move touches, bit 1 a pair inside the band somewhere,
// bits 8+j / 16+j certain overlap with the chain of move j at its old / trial position.
template <int NB, int NWARP>
__device__ __forceinline__ unsigned team_test(const WCtx &w, const Filt &f, const int *wmeta, int *lst, float4 *buf, const float *rtw,
                                              bool active, int k, int ln, int ichain, int b0, bool intra, float Rt,
                                              unsigned &pairs)
{
    const int nb = NB ? NB : w.nb, nc = w.nc, m = nb >> 1, f4 = W_F4(w), c4 = w.w4 + nb * k;
    int early[NWARP > 1 ? NWARP - 1 : 1];
#pragma unroll
    for (int j = 0; j < NWARP - 1; ++j) early[j] = j < k ? (wmeta[j] & 0xffff) : -1;
#if !HANS_TEAM_BEAD_PRUNE
    float4 tr[NB ? NB : 1];
    if (NB) {
#pragma unroll
        for (int b = 0; b < NB; ++b) tr[b] = SM4[c4 + b];
    }
#endif
    const float4 ct = SM4[c4 + m];
    bool ov = false, amb = false;
    // TEAM_WAVES waves of 32 chain spheres at a time (independent tests in flight), ONE compaction, one pass over
    // the beads of the survivors
#pragma unroll 1
    for (int cb = 0; active && cb < nc; cb += TEAM_LIST) {
        bool near[TEAM_WAVES];
#pragma unroll
        for (int u = 0; u < TEAM_WAVES; ++u) {
            const int c = cb + 32 * u + ln;
            near[u] = false;
            if (c < nc && c != ichain) {
                bool touched = false;
#pragma unroll
                for (int j = 0; j < NWARP - 1; ++j) touched |= (c == early[j]);
                near[u] = !touched && spheres_near(f, SM4[f4 + c * nb + m], ct, Rt + SMF[w.rad + c] + 1.001f);
                pairs += 1u;
            }
        }
        int cnt = 0;
#pragma unroll
        for (int u = 0; u < TEAM_WAVES; ++u) {
            if (cb + 32 * u < nc) {
                const unsigned mask = __ballot_sync(FULL, near[u]);
                if (near[u]) lst[cnt + __popc(mask & ((1u << ln) - 1u))] = cb + 32 * u + ln;
                cnt += __popc(mask);
            }
        }
        __syncwarp();
        const int items = cnt * nb;
#if HANS_TEAM_BEAD_PRUNE
        // bead level, two steps: which beads of the near chains lie inside the trial chain's bounding sphere (+ sigma)
        // -- one test per bead; then (bead, trial bead) pairs of the survivors spread over the lanes
#pragma unroll 1
        for (int e0 = 0; e0 < items; e0 += 32) {
            const int e = e0 + ln;
            bool in = false;
            float4 s = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (e < items) {
                const int ci = NB ? e / (NB ? NB : 1) : div_nb(w, e);
                s = SM4[f4 + lst[ci] * nb + (e - ci * nb)];
                in = spheres_near(f, s, ct, Rt + 1.001f);
                pairs += 1u;
            }
            const unsigned mask = __ballot_sync(FULL, in);
            if (in) buf[__popc(mask & ((1u << ln) - 1u))] = s;
            __syncwarp();
            const int npair = __popc(mask) * nb;
#pragma unroll 1
            for (int t = ln; t < npair; t += 32) {
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
#else
#pragma unroll 1
        for (int e = ln; e < items; e += 32) {
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
#pragma unroll 1
                for (int b = b0; b < nb; ++b) {
                    const float4 t = SM4[c4 + b];
                    const float r2 = r2_f32(f, s.x - t.x, s.y - t.y, s.z - t.z);
                    ov |= (r2 < f.lo);
                    amb |= !(r2 > f.hi);
                }
            }
            pairs += (unsigned)(nb - b0);
        }
#endif
        if (cb + TEAM_LIST < nc) __syncwarp(); // the list is rewritten by the next chunk
    }
    if (active && intra) { // dihedral: pairs (j, b), b >= b0, of the trial chain itself beyond the exclusion
        const uchar2 *tab = W_ITAB(w);
#pragma unroll 1
        for (int e = ln; e < w.npc; e += 32) {
            const uchar2 jb = tab[e];
            if ((int)jb.y >= b0) {
                const float4 p = SM4[c4 + jb.x], q = SM4[c4 + jb.y];
                const float r2 = r2_f32(f, q.x - p.x, q.y - p.y, q.z - p.z);
                ov |= (r2 < f.lo);
                amb |= !(r2 > f.hi);
                pairs += 1;
            }
        }
    }
    // the chains of the earlier moves of the window, at their old position (F4, RAD) and at their trial position
    // (their w4 slot, rtw): bounding spheres first, lane j for the old and lane 8 + j for the new position of move j,
    // bead level only for the ones that are near
    __syncthreads(); // the trial chains of the other warps (written before their own tests) are visible from here on
    unsigned oldb = 0, newb = 0;
    if (active) {
        bool nr = false;
        const int j = ln & 7;
        if (ln < 16 && j < k) {
            const int cj = wmeta[j] & 0xffff;
            const bool isnew = ln >= 8;
            const float4 cc = isnew ? SM4[w.w4 + nb * j + m] : SM4[f4 + cj * nb + m];
            const float rr = isnew ? rtw[j] : SMF[w.rad + cj];
            nr = spheres_near(f, cc, ct, Rt + rr + 1.001f);
            pairs += 1u;
        }
        unsigned nm = __ballot_sync(FULL, nr);
        while (nm) {
            const int bit = __ffs(nm) - 1;
            nm &= nm - 1u;
            const int jj = bit & 7;
            const int s4 = bit >= 8 ? w.w4 + nb * jj : f4 + (wmeta[jj] & 0xffff) * nb;
            bool o = false;
#pragma unroll 1
            for (int t = ln; t < nb * nb; t += 32) { // (rare: the spheres of two moves of one window touch)
                const int e = NB ? t / (NB ? NB : 1) : div_nb(w, t);
                const int b = t - e * nb;
                const float4 s = SM4[s4 + e], tb = SM4[c4 + b];
                const float r2 = r2_f32(f, s.x - tb.x, s.y - tb.y, s.z - tb.z);
                const bool moved = b >= b0;
                o |= moved && (r2 < f.lo);
                amb |= moved && !(r2 > f.hi);
                pairs += moved ? 1u : 0u;
            }
            o = __any_sync(FULL, o);
            if (bit >= 8) newb |= (o ? 1u : 0u) << jj; else oldb |= (o ? 1u : 0u) << jj;
        }
    }
    return __reduce_or_sync(FULL, (ov ? 1u : 0u) | (amb ? 2u : 0u) | (oldb << 8) | (newb << 16));
}

template <int NB, int NWARP, bool REPLAY>
__global__ void __launch_bounds__(32 * NWARP, 16 / NWARP) walk_team_kernel(const WalkArgs a)
{
    constexpr int T = 32 * NWARP;
    static_assert(NWARP >= 2 && NWARP <= 8, "result words carry 8 bits per earlier move");
    const WCtx w = setup_ctx_g<T, 1>(a.cfg, NWARP, team_tail_bytes(NWARP));
    const int wid = threadIdx.x >> 5, ln = threadIdx.x & 31;
    const int nb = NB ? NB : w.nb;
    unsigned *res = reinterpret_cast<unsigned *>(smem) + a.team_res; // (walker + window area) / 4, from the launcher
    float *rtw = reinterpret_cast<float *>(res) + 24;
    unsigned *cnt = res + 8; // 12 counters in shared memory: twelve registers less in a kernel held at 128
    int *meta = reinterpret_cast<int *>(res) + 32, *info = meta + T + 8;
    int *lst = reinterpret_cast<int *>(res) + 48 + 2 * T + TEAM_LIST * wid; // this warp's near list
    float4 *buf = reinterpret_cast<float4 *>(res + 48 + 2 * T + TEAM_LIST * NWARP) + 32 * wid;
    const double limit = a.d_limit ? *a.d_limit : a.limit;
    for (int k = blockIdx.x; k < a.n_active; k += gridDim.x) {
        const int wi = a.ids ? a.ids[k] : k;
        load_walker<T>(w, a.H, a.R, wi, true);
        Filt f = load_filt(W_SF(w));
        const unsigned long long ctr0 = REPLAY ? 0ull : a.ctr[wi];
        if (threadIdx.x < 12) cnt[threadIdx.x] = 0u; // attempted / accepted per move type (ordered by the first barrier of a batch)
        unsigned pairs = 0;
        unsigned long long pairs_total = 0;
        PropRec *ring = W_RING(w);
        for (int m0 = 0; m0 < a.nmoves; m0 += T) {
            const int nbatch = a.nmoves - m0 < T ? a.nmoves - m0 : T;
            int my_type = -1, my_meta = -1;
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
                // a single-chain move with a well-formed proposal
                const bool plain = (my_type == HANS_TRANS || my_type == HANS_ROT ||
                                    (my_type == HANS_DIH && nb >= 4 && slot->idx0 >= 3 && slot->idx0 <= nb - 1)) &&
                                   (unsigned)slot->ichain < (unsigned)w.nc;
                if (plain) my_meta = slot->ichain | ((slot->u_acc < 1.0) ? (1 << 30) : 0); // MCNS.py:372 with boltz = 1
            }
            meta[w.lane] = my_meta;
            if (w.lane < 8) meta[T + w.lane] = -1;
            __syncthreads();
            {   // the run of plain moves of DISTINCT chains that starts at my move (at most NWARP), once per batch
                int n = 0;
                unsigned um = 0;
                int ch[NWARP];
                bool live = true;
#pragma unroll
                for (int j = 0; j < NWARP; ++j) {
                    const int mj = meta[w.lane + j];
                    ch[j] = mj & 0xffff;
                    bool ok = live && mj >= 0;
#pragma unroll
                    for (int i = 0; i < j; ++i) ok = ok && (ch[i] != ch[j]);
                    live = ok;
                    n += ok ? 1 : 0;
                    um |= (ok && (mj >> 30)) ? (1u << j) : 0u;
                }
                info[w.lane] = n | (int)(um << 8);
            }
            __syncthreads();
            int my_acc = 0, my_reason = 0;
            double my_boltz = 0.0;
            for (int q = 0; q < nbatch;) {
                const int inf = info[q];
                const int nw = inf & 0xff;          // moves of this window
                const unsigned ucm = (unsigned)inf >> 8; // bit i = move q + i commits if it is accepted
                if (nw == 0) { // box move or malformed proposal: the whole CTA on one move
                    const int type = ring[q].type;
                    int accepted, reason;
                    double boltz;
                    if (type >= HANS_TRANS && type <= HANS_DIH) {
                        const Dec d = move_chain<T>(w, ring + q, HANS_MODE_FUSED, pairs);
                        accepted = d.accepted; reason = d.reason; boltz = d.boltz;
                    } else {
                        const DecP d = move_box<T>(w, q, limit, HANS_MODE_FUSED);
                        accepted = d.accepted; reason = d.reason; boltz = d.boltz; pairs += d.pairs;
                        f = load_filt(W_SF(w)); // the cell may have changed (also on a reverted shear / stretch)
                    }
                    if (w.lane == q) { my_acc = accepted; my_reason = reason; my_boltz = boltz; }
                    ++q;
                    continue;
                }
                // ---- A1: warp k builds the trial chain of move k ---------------------------------
                int ichain = 0, b0 = 0;
                bool intra = false;
                const int cnk = w.wc + 3 * nb * wid, c4k = w.w4 + nb * wid;
                float Rt = 0.0f;
                const bool active = wid < nw;
                if (active) {
                    ichain = meta[q + wid] & 0xffff;
                    make_trial(w, ring + q + wid, ln, 32, cnk, c4k, b0, intra);
                    __syncwarp();
                    Rt = intra ? chain_radius_warp(cnk, nb) : SMF[w.rad + ichain]; // translations and rotations are rigid
                    if (ln == 0) rtw[wid] = Rt;
                }
                // ---- A2: warp k tests it (one barrier inside, before the other warps' trial chains are read) ----
                {
                    const unsigned r = team_test<NB, NWARP>(w, f, meta + q, lst, buf, rtw, active, wid, ln, ichain, b0, intra, Rt, pairs);
                    if (active && ln == 0) res[wid] = r;
                }
                __syncthreads();
                // ---- B: settle in order (uniform over the CTA, branch free) ------------------------
                unsigned committed = 0, accepted = 0;
                int ndone = nw;
                {
                    unsigned r[8];
                    *reinterpret_cast<uint4 *>(r) = *reinterpret_cast<const uint4 *>(res);
                    if (NWARP > 4) *reinterpret_cast<uint4 *>(r + 4) = *reinterpret_cast<const uint4 *>(res + 4);
                    unsigned cert = 0, other = 0; // per move: certain overlap; anything else (band, earlier moves)
#pragma unroll
                    for (int i = 0; i < NWARP; ++i) {
                        const unsigned ri = i < nw ? r[i] : 0u;
                        cert |= (ri & 1u) << i;
                        other |= ri >> 1;
                    }
                    if (other == 0u) { // the usual case: no move of the window is near an earlier one's chain
                        accepted = ~cert & ((1u << nw) - 1u);
                        committed = accepted & ucm;
                    } else {
                        bool live = true;
#pragma unroll
                        for (int i = 0; i < NWARP; ++i) {
                            const unsigned ri = r[i];
                            const bool in = live && i < nw;
                            const bool certain = (ri & 1u) != 0u;
                            const bool cut = in && !certain && (ri & 2u) != 0u; // a pair inside the band: float64 decides, serially
                            ndone = cut ? i : ndone;
                            live = live && !cut;
                            const bool ov = certain || ((((ri >> 16) & committed) | ((ri >> 8) & ~committed)) & ((1u << i) - 1u)) != 0u;
                            const unsigned bit = (in && !cut && !ov) ? (1u << i) : 0u;
                            accepted |= bit;
                            committed |= bit & ucm;
                        }
                    }
                }
                // ---- C: commit -----------------------------------------------------------------------
                if (wid < ndone && ((committed >> wid) & 1u)) {
                    const int P = W_P(w) + 3 * nb * ichain, f4c = W_F4(w) + nb * ichain;
                    if (NB && NB % 2 == 0 && 5 * NB / 2 <= 32) {
                        // beads (3 NB doubles = 3 NB / 2 sixteen-byte words; P and cnk are even) and shadows: one pass
                        constexpr int n2 = 3 * NB / 2;
                        if (ln < n2 + NB) {
                            const int src = ln < n2 ? (cnk >> 1) + ln : c4k + (ln - n2);
                            const int dst = ln < n2 ? (P >> 1) + ln : f4c + (ln - n2);
                            SM4[dst] = SM4[src];
                        }
                    } else {
                        for (int e = ln; e < 3 * nb; e += 32) smem[P + e] = smem[cnk + e];
                        for (int e = ln; e < nb; e += 32) SM4[f4c + e] = SM4[c4k + e];
                    }
                    if (intra && ln == 0) SMF[w.rad + ichain] = Rt;
                }
                {
                    const int d = w.lane - q;
                    if (d >= 0 && d < ndone) {
                        my_acc = (int)((committed >> d) & 1u); // accepted = no overlap AND u_acc < 1 (MCNS.py:372)
                        my_reason = my_acc ? HANS_ACCEPT : HANS_REJ_OVERLAP;
                        my_boltz = my_acc ? 1.0 : 0.0;
                    }
                }
                __syncthreads();
                q += ndone;
                if (ndone < nw) { // move q had a pair inside the band: the cooperative path settles it in float64
                    const Dec d = move_chain<T>(w, ring + q, HANS_MODE_FUSED, pairs);
                    if (w.lane == q) { my_acc = d.accepted; my_reason = d.reason; my_boltz = d.boltz; }
                    ++q;
                }
            }
            pairs_total += pairs;
            pairs = 0;
            if (my_type >= 0) {
                if (my_type < 6) {
                    atomicAdd(cnt + my_type, 1u);
                    if (my_acc) atomicAdd(cnt + 6 + my_type, 1u);
                }
                if (REPLAY && a.out) {
                    hans_decision *o = a.out + (size_t)k * a.nmoves + m0 + w.lane;
                    o->accepted = my_acc; o->reason = my_reason; o->boltz = my_boltz;
                }
            }
            __syncthreads(); // the ring, meta and info are rewritten by the next batch
        }
        store_walker<T>(w, a.H, a.R, wi);
        if (a.attempted && threadIdx.x < 6) { // (the batch loop ends with a barrier)
            atomicAdd(a.attempted + (size_t)k * 6 + threadIdx.x, (unsigned long long)cnt[threadIdx.x]);
            atomicAdd(a.accepted + (size_t)k * 6 + threadIdx.x, (unsigned long long)cnt[6 + threadIdx.x]);
        }
        if (w.lane == 0) {
            if (!REPLAY) a.ctr[wi] = ctr0 + (unsigned long long)a.nmoves;
            else if (a.ctr) a.ctr[wi] += a.ctr_add; // proposals came from hans_gen_proposals
            if (a.vol) a.vol[wi] = fabs(det3(smem + W_SH(w)));
        }
        if (a.pair_tests) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) pairs_total += __shfl_down_sync(FULL, pairs_total, o);
            if (ln == 0 && pairs_total) atomicAdd(a.pair_tests, pairs_total);
        }
        __syncthreads();
    }
}

} // namespace hb
