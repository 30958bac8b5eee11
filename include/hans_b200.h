/*
 * hans_b200.h -- C ABI of libhans_b200.so, the B200 (sm_100a) replacement for the box / chain /
 * overlap entry points of hs_alkane that hans' nested-sampling walker engine calls
 * (import at /root/reference/NesSa/MCNS.py:2; call sites listed per function below).
 *
 * Boundary rules
 *   - plain C: pointers, sizes, POD structs; no torch / C++ types.
 *   - every pointer named d_* is a DEVICE pointer (the Python host layer allocates them as torch
 *     CUDA tensors and passes .data_ptr()); pointers without the prefix are host pointers.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); all work is
 *     enqueued asynchronously on it, nothing synchronises unless stated.
 *   - every entry returns 0 on success, a negative HANS_E* code otherwise (never exits);
 *     hans_last_error() gives the text.
 *   - there is NO CPU fallback: without a CUDA device every compute entry returns HANS_ECUDA.
 *
 * State layout (float64, identical to what the reference stores in restart.hdf5,
 * NSio.py:86-95): for walker w of a pool,
 *     d_H[w*9 .. w*9+8]            3x3 cell, rows = cell vectors (MCNS.py:112-113)
 *     d_R[(w*N + i)*3 .. +2]       Cartesian bead i = ichain*nbeads + ibead, N = nchains*nbeads
 *     d_ctr[w]                     number of trial moves this walker has consumed (Philox counter)
 *     d_vol[w]                     |det H| as of the walker's last walk
 * Indices are 0-based in this ABI; the Python shim converts from the reference's 1-based ibox /
 * ichain.
 */
#ifndef HANS_B200_H
#define HANS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HANS_ABI_VERSION 1

enum { HANS_OK = 0, HANS_EINVAL = -1, HANS_ECUDA = -2, HANS_ELIMIT = -3 };

/* move types, order of move_ratio (MCNS.py:25-27) */
enum { HANS_VOL = 0, HANS_TRANS = 1, HANS_ROT = 2, HANS_DIH = 3, HANS_SHEAR = 4, HANS_STRETCH = 5 };

/* decision reasons */
enum {
    HANS_ACCEPT = 0,
    HANS_REJ_OVERLAP = 1,
    HANS_REJ_CEILING = 2,
    HANS_REJ_BOLTZ = 3,
    HANS_REJ_ASPECT = 4,
    HANS_REJ_ANGLE = 5,
    HANS_REJ_INVALID = 6
};

enum {
    HANS_MODE_FUSED = 0,  /* MC_run semantics (MCNS.py:349-380): decide, revert on reject */
    HANS_MODE_PROPOSE = 1 /* hs_alkane entry-point semantics: leave trial state, return boltz */
};

#define HANS_MAX_BEADS 64 /* beads per chain */

/* OR-ed into threads_per_walker (0, 32, 64, 128, 256): the cooperative kernel with the PER-WALKER CELL LIST -- bins in
 * fractional space at least 1.001 sigma wide in every direction, overlap candidates of a bead = the beads of the <= 27
 * bins around it (the role of hs_alkane's link cells, bypassed by the reference at MCNS.py:603).  Same decisions and
 * trajectories as every other kernel, bit for bit.  With threads_per_walker = 0 the library uses it by itself for walkers
 * that do not fit shared memory (state in global memory); for the BASELINE shapes the sphere-pruning kernels measured faster
 * at every density, so it stays an explicit choice there. */
#define HANS_TPW_CELLS 4096

/* Model + move parameters.  Replaces hs_alkane's module state set through
 * alkane_set_nchains/nbeads/bondlength/bondangle, alkane_set_d{v,r,t,h}_max (MCNS.py:599-606,
 * mpihans.py:118-133) and the MC_run arguments (MCNS.py:276-277). */
typedef struct hans_cfg {
    int32_t nchains;
    int32_t nbeads;
    int32_t nexclude;   /* intra-chain pairs tested iff |i-j| > nexclude (default 3) */
    int32_t allow_flip; /* alkane_bond_rotate(.., allow_flip), MCNS.py:333 */
    double dv_max, dr_max, dt_max, dh_max, dshear, dstretch;
    double min_ar;      /* MCNS.py:194,243 */
    double cos_min_ang; /* cos(min_ang in radians), MCNS.py:190-196 */
    double pressure;    /* MCNS.py:277 */
    double move_cum[6]; /* cumsum(move_ratio)/sum, MCNS.py:294 */
    double bondlength;  /* alkane_set_bondlength, MCNS.py:605 */
    double cos_bondangle, sin_bondangle; /* alkane_set_bondangle, MCNS.py:606 */
} hans_cfg;

/* One explicit trial move (replay mode). */
typedef struct hans_proposal {
    int32_t type;
    int32_t ichain;
    int32_t idx0; /* dih: first rotated bead | shear: k | stretch: i */
    int32_t idx1; /* dih: flip | stretch: j */
    double p[4];  /* vol: dV | trans: dr[3] | rot: quaternion wxyz | dih: angle | shear: rv1,rv2 |
                     stretch: rv.  (Fields a move type does not use are ignored on input; the walk kernels use p[1], p[2] of
                     their shared-memory copy of a dihedral record for sin / cos of the half angle.) */
    double u_acc; /* the uniform draw of MCNS.py:351,372: accepted iff u_acc < boltz.  Generated values lie in [0, 1); an
                     explicit value >= 1 makes a single-chain move a rejection (reason HANS_REJ_OVERLAP, as the reference's
                     else-branch would report it) */
    double reserved;
} hans_proposal;

typedef struct hans_decision {
    int32_t accepted;
    int32_t reason;
    double boltz;
} hans_decision;

int hans_abi_version(void);
const char *hans_last_error(void);
/* number of CUDA devices visible (0 if none / no driver); never fails */
int hans_device_count(void);
/* shared memory bytes one walker needs; lets the host pick threads_per_walker */
int64_t hans_walker_smem_bytes(const hans_cfg *cfg);
/* MCNS.py:296-301 */
int hans_moves_per_sweep(int nbeads, int nchains);
/* threads per walker of the cooperative kernel for a shape (32 = one warp per walker, four walkers
 * per CTA; 64 / 128 = one CTA per walker) */
int hans_default_threads_per_walker(const hans_cfg *cfg);
/* window size the warp-per-walker kernel uses for a shape when threads_per_walker = 0 (1 = one
 * move at a time, 8 = speculative windows); 0 = the shape runs on the cooperative kernel */
int hans_default_window(const hans_cfg *cfg);
/* warps per walker of the warp-per-move kernel (one CTA per walker, a window of consecutive chain moves
 * of distinct chains in flight, one warp each) when threads_per_walker = 0 picks it for a shape
 * (large walkers with many chains); 0 = the shape runs on one of the other kernels */
int hans_default_team(const hans_cfg *cfg);

/* ---- the fused hot path ---------------------------------------------------------------------
 * MC_run (NesSa/MCNS.py:276-392) for n_active walkers at once: sweeps * moves_per_sweep trial
 * moves each, proposals drawn from Philox4x32-10 keyed by (seed; d_ctr[w], gid_base + w).
 * d_ids[n_active]     walkers to walk (NULL = walkers 0..n_active-1)
 * d_volume_limit      device scalar with the volume ceiling (NULL = use volume_limit)
 * d_attempted/d_accepted  [n_active][6] counters, ACCUMULATED into (may be NULL)
 * d_vol               [pool] volumes, entry w rewritten for every walked w (may be NULL)
 * d_pair_tests        device scalar, executed pair tests accumulated (may be NULL)
 * threads_per_walker  32, 64, 128, 256: cooperative kernel, the walker's threads share every move
 *                     (32 = one warp per walker, else one CTA per walker);
 *                     1, 8: warp-per-walker kernel specialised on nbeads (nbeads in
 *                     {1,2,4,6,8,12}; up to 128 beads, or chains of up to 64 chains and fewer than
 *                     256 beads); 1 = one move at a time, 8 = speculative windows of up to 8
 *                     consecutive single-chain moves, settled in order (<= 128 beads; same
 *                     decisions and trajectory, bit for bit);
 *                     24, 28: warp-per-move kernel for large chain systems (nbeads > 1): one CTA of
 *                     4 / 8 warps per walker, up to 4 / 8 consecutive single-chain moves of distinct
 *                     chains evaluated at once, one warp each, and settled in order (same decisions
 *                     and trajectory, bit for bit); box moves are shared by the whole CTA;
 *                     any of 0, 32, 64, 128, 256 | HANS_TPW_CELLS: cooperative kernel with the per-walker cell list;
 *                     0 = the measured default per shape: hans_default_window if non-zero, else
 *                     hans_default_team if non-zero, else hans_default_threads_per_walker (walkers beyond shared memory:
 *                     state in global memory + the cell list).
 *                     Every variant produces the same trajectory; the choice is speed only.
 * Replaces: the Python loop MCNS.py:304-383 and every alk.* call inside it. */
int hans_mc_run_batch(const hans_cfg *cfg, double *d_H, double *d_R, uint64_t *d_ctr,
                      const int32_t *d_ids, int n_active, int sweeps,
                      const double *d_volume_limit, double volume_limit, uint64_t seed,
                      uint32_t gid_base, uint64_t *d_attempted, uint64_t *d_accepted,
                      double *d_vol, uint64_t *d_pair_tests, int threads_per_walker, void *stream);

/* The same walk with the proposals generated ahead of it: proposals do not depend on the state, so
 * hans_gen_proposals (throughput-bound: Philox, Box-Muller, quaternions) fills d_workspace at full
 * occupancy and the latency-bound walk kernel reads 64 bytes per move instead of computing them on
 * its critical path.  Same trajectory, bit for bit.  The workspace holds n_active * 64 bytes per
 * move; if it is smaller than the walk, the walk is cut into chunks of what fits (generate, walk,
 * generate, ...);
 * d_workspace = NULL falls back to generation inside the walk kernel (= hans_mc_run_batch). */
int hans_mc_run_batch_ws(const hans_cfg *cfg, double *d_H, double *d_R, uint64_t *d_ctr,
                         const int32_t *d_ids, int n_active, int sweeps,
                         const double *d_volume_limit, double volume_limit, uint64_t seed,
                         uint32_t gid_base, uint64_t *d_attempted, uint64_t *d_accepted,
                         double *d_vol, uint64_t *d_pair_tests, int threads_per_walker,
                         void *d_workspace, uint64_t workspace_bytes, void *stream);

/* kernel launches hans_mc_run_batch_ws makes for a walk of nmoves moves per walker (for launch accounting) */
int hans_mc_run_launches(int n_active, int nmoves, int have_workspace, uint64_t workspace_bytes);

/* d_props[k][m] = proposal number d_ctr[walker k] + m of walker k = d_ids[k], m < nmoves: exactly
 * the proposals hans_mc_run_batch would draw (four Philox4x32-10 blocks per move
 * keyed by (seed; counter, gid_base + walker)).  Does not advance d_ctr. */
int hans_gen_proposals(const hans_cfg *cfg, const uint64_t *d_ctr, const int32_t *d_ids, int n_active,
                       int nmoves, uint64_t seed, uint32_t gid_base, hans_proposal *d_props,
                       void *stream);

/* ---- replay: explicit proposals in, decisions out (bit-exact parity entry) -------------------
 * d_props[n_active][nmoves], d_out[n_active][nmoves].  mode = HANS_MODE_FUSED | HANS_MODE_PROPOSE.
 * With HANS_MODE_PROPOSE and nmoves = 1 this IS alk.alkane_translate_chain / alkane_rotate_chain /
 * alkane_bond_rotate / alkane_box_resize(..,0) / box_shear_step / box_stretch_step
 * (MCNS.py:318-343): the trial state is left in place and boltz is returned. */
int hans_replay(const hans_cfg *cfg, double *d_H, double *d_R, const int32_t *d_ids, int n_active,
                const hans_proposal *d_props, int nmoves, double volume_limit, int mode,
                hans_decision *d_out, int threads_per_walker, void *stream);

/* ---- predicates -----------------------------------------------------------------------------
 * alk.alkane_check_chain_overlap (MCNS.py:198,247,543) for n walkers: d_out[k] = 1 if any two
 * beads of different chains (and, unless inter_only, non-excluded beads of one chain) are closer
 * than sigma = 1 under the minimum image convention. */
int hans_check_overlap(const hans_cfg *cfg, const double *d_H, const double *d_R,
                       const int32_t *d_ids, int n, int inter_only, int32_t *d_out,
                       int threads_per_walker, void *stream);

/* alk.box_compute_volume, alk.box_min_aspect_ratio (MCNS.py:97-118,194) and the largest |cos|
 * between cell vectors (min_angle = arccos of it, MCNS.py:120-133) for n cells; any output may be
 * NULL. */
int hans_cell_metrics(const double *d_H, int n, double *d_vol, double *d_min_ar, double *d_max_cos,
                      void *stream);

/* alk.alkane_change_box(ibox, dH) (MCNS.py:186,238,368): H += dH, chains follow rigidly.
 * d_dH[n][9]. */
int hans_change_box(const hans_cfg *cfg, double *d_H, double *d_R, const int32_t *d_ids, int n,
                    const double *d_dH, void *stream);

/* The cell change a shear / stretch proposal makes, as the reference's Python builds it: delta_H of
 * box_shear_step (MCNS.py:148-184) and box_stretch_step (MCNS.py:220-234), the value those functions
 * return and MC_run reverts with (MCNS.py:366-368).  d_props[n] (one per walker d_ids[k]), d_dH[n][9];
 * proposals of other types give zero.  Does not touch the walkers. */
int hans_shape_delta(const double *d_H, const int32_t *d_ids, int n, const hans_proposal *d_props,
                     double *d_dH, void *stream);

/* ---- initial configurations -------------------------------------------------------------------
 * create_initial_configs + populate_boxes (MCNS.py:628-644): cells must already be set in d_H;
 * grows all chains of n walkers by random insertion (retry until no overlap, at most
 * max_attempts per chain).  d_status[k] = attempts used, or -1 if a chain did not fit. */
int hans_grow(const hans_cfg *cfg, const double *d_H, double *d_R, int n, uint64_t seed,
              uint32_t gid_base, int max_attempts, int32_t *d_status, void *stream);
/* one attempt for one chain of one walker: alk.alkane_grow_chain(ichain, ibox, 1) (MCNS.py:642);
 * d_status[0] = 1 if the chain was placed, -1 if it overlapped (ifail). */
int hans_grow_chain(const hans_cfg *cfg, const double *d_H, double *d_R, int walker, int ichain,
                    uint64_t seed, uint32_t gid, uint32_t attempt, int32_t *d_status, void *stream);

/* ---- nested-sampling iteration helpers (mpihans.py:210-245) -----------------------------------
 * Local MAXLOC over d_vol[0..n) (mpihans.py:211-212): writes d_out = {max volume, (double)index};
 * ties resolve to the lowest index, as list.index does. */
int hans_argmax(const double *d_vol, int n, double *d_out2, void *stream);

/* Clone (mpihans.py:226-237, MCNS.py:550-580): copy cell + beads (+ volume) of walker src of the
 * source pool over walker dst of the destination pool.  The source may be a staging buffer
 * received from another GPU (then src = 0).  d_dst_index/d_src_index are optional DEVICE scalars
 * (int32) that override dst/src, so the copy can be enqueued before the host knows the argmax.
 * If d_enable is non-NULL the copy happens only when *d_enable != 0. */
int hans_clone(const hans_cfg *cfg, const double *d_srcH, const double *d_srcR,
               const double *d_srcvol, int src, const int32_t *d_src_index, double *d_dstH,
               double *d_dstR, double *d_dstvol, int dst, const int32_t *d_dst_index,
               const int32_t *d_enable, void *stream);

/* ---- one nested-sampling iteration without host round trips (mpihans.py:210-245) ---------------
 * The host enqueues, per iteration:  hans_ns_pack -> [all-gather of the per-rank records over
 * NCCL when world > 1] -> hans_ns_resolve -> hans_mc_run_batch(d_ids, d_volume_limit = d_limit).
 * Everything that varies per iteration (iteration number, clone source, active walkers, ceiling)
 * lives in device memory, so a chunk of iterations can be enqueued (or graph-captured) without
 * synchronising.
 *
 * Walkers are sharded nlocal per rank (mpihans.py:54,185); a rank may host several "virtual ranks"
 * of nw_per_vrank walkers each, and -- exactly like one MPI rank of the reference,
 * mpihans.py:232-241 -- every virtual rank walks ONE walker per iteration: the freshly cloned
 * max-volume slot if it owns it, else a uniformly random one of its walkers.
 *
 * Record layout (doubles): [0] local max volume, [1] its local index, [2] volume of the staged
 * clone source, [3..11] its cell, [12..12+3N) its beads.  The clone source is
 * g = Philox(seed; iteration) mod (nlocal*world) on every rank alike (replaces the randint +
 * bcast of mpihans.py:219-223); only the rank that owns it stages it (mpihans.py:226-229).
 */
typedef struct hans_ns_args {
    int32_t N;             /* beads per walker */
    int32_t nlocal;        /* walkers resident on this rank */
    int32_t rank, world;
    int32_t nw_per_vrank;  /* the reference's nwalkers (per MPI rank); nlocal % nw_per_vrank == 0 */
    int32_t traj_interval; /* mpihans.py:38; 0 = never stage trajectory frames */
    int32_t traj_slots;
    int32_t log_cap;       /* entries in d_log (ring) */
    uint64_t seed;
    double *d_H, *d_R, *d_vol;
    int64_t *d_iter;          /* device iteration counter; incremented by hans_ns_resolve */
    double *d_rec;            /* this rank's record, 12 + 3N doubles */
    const double *d_gathered; /* world records (== d_rec when world == 1) */
    double *d_limit;          /* out: the new volume ceiling = global max volume (mpihans.py:215) */
    double *d_log;            /* [log_cap][3]: max volume, owner rank, owner index; slot = iter % log_cap */
    int32_t *d_ids;           /* out: [nlocal / nw_per_vrank] walkers to walk this iteration */
    double *d_traj;           /* [traj_slots][12+3N]: the max walker BEFORE it is overwritten
                                 (mpihans.py:232-237), staged when iter % traj_interval == 0 */
    int64_t *d_traj_iter;     /* [traj_slots] iteration staged in each slot (-1 = empty) */
} hans_ns_args;

/* local MAXLOC (mpihans.py:211-212) + stage the clone source if this rank owns it */
int hans_ns_pack(const hans_ns_args *a, void *stream);
/* global MAXLOC over the gathered records (mpihans.py:215; ties -> lowest rank, then index), new
 * ceiling, volumes log, trajectory staging, clone into the max slot (mpihans.py:232-237), active
 * walker list (mpihans.py:236-239), iteration counter += 1 */
int hans_ns_resolve(const hans_ns_args *a, void *stream);

/* ---- the same exchange over NVLink peer memory (one node, CUDA IPC) -----------------------------------
 * hans_ns_exchange_peer = hans_ns_pack + all-gather + hans_ns_resolve in ONE kernel per rank: every rank
 * stores its record straight into a buffer of every other rank and raises a flag there; each rank
 * waits for the world flags in its own buffer and resolves from its own copy (no collective library,
 * no kernel boundary; protocol in hans_kernels.cu).  Set-up, once: every rank calls hans_peer_alloc
 * (hans_ns_peer_bytes(N, world) bytes, zeroed) and publishes the 64-byte handle to the other ranks of
 * the node; each rank opens the others' handles (hans_peer_open) and passes the world base pointers --
 * its own at index rank -- as the HOST array peer_bufs.  a->d_gathered is not used (but must be non-NULL).
 * Opt-in in the host layer (HANS_NS_PEER=1) until it has been timed on several GPUs; the NCCL all_gather
 * path is the default. */
#define HANS_PEER_MAX 16
uint64_t hans_ns_peer_bytes(int N, int world);
int hans_peer_alloc(uint64_t bytes, void **d_ptr, unsigned char *handle64);
int hans_peer_open(const unsigned char *handle64, void **d_ptr);
int hans_peer_close(void *d_ptr);
int hans_peer_free(void *d_ptr);
int hans_ns_exchange_peer(const hans_ns_args *a, void *const *peer_bufs, void *stream);
/* bound the wait for the other ranks' flags (seconds of SM clock; 0 = wait for ever, the default, like a
 * collective).  A wait that expires is counted in hans_diag_counters out4[2] and the iteration's result is void. */
int hans_ns_peer_timeout(double seconds);

/* ---- order parameters (the reference's offline nematic.py) ---------------------------
 * For n configurations: d_p2[k] = nematic order parameter P2, the largest eigenvalue of
 * Q = mean over chains of (3 u u^T - I)/2 with u the unit minimum-image end-to-end vector
 * (nematic.py:83-137; 0 for nbeads = 1); d_trans[k] = the translational order parameter
 * |sum_chains exp(2 pi i (s_x+s_y+s_z))| / nchains over the fractional chain centres as the
 * reference takes them (mean of beads 0..nbeads-2, nematic.py:30-72).  Either output may be NULL. */
int hans_order_params(const hans_cfg *cfg, const double *d_H, const double *d_R, int n, double *d_p2,
                      double *d_trans, void *stream);

/* Diagnostics (synchronises): out4[0] = trial chains re-tested in float64 because a pair fell into
 * the float32 filter's band, out4[1] = float64 all-pairs passes of box moves, out4[2] = expired waits of
 * hans_ns_exchange_peer; reset != 0 zeroes them. */
int hans_diag_counters(uint64_t *out4, int reset);

/* FP32 FFMA issue-rate probe used as the roofline denominator in bench.py: runs `iters`
 * dependent-chain-free FFMAs per thread on the whole chip, returns achieved TFLOP/s. */
int hans_ffma_peak(int iters, double *tflops, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* HANS_B200_H */
