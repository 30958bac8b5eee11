/*
 * hans_oracle.h -- CPU oracle for the hans nested-sampling Monte Carlo walker path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under hans_b200/ may include, link, import or execute this;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * PARITY: PINNED to everything the reference itself computes, UNPINNED for hs_alkane's internals.
 *   (1) The parts the reference spells out in Python -- the MC_run control flow (NesSa/MCNS.py:276-392: move and chain
 *       selection, draw order and count, accept / revert / ceiling logic, rate bookkeeping), the shear and stretch moves
 *       (MCNS.py:135-252: delta_H, rejection order), the cell-shape predicates (MCNS.py:97-133), moves_per_sweep
 *       (MCNS.py:296-301), the step-size rule (MCNS.py:687-716), default_move_ratio (MCNS.py:610-626) -- are pinned by
 *       EXECUTING the unmodified /root/reference/NesSa/MCNS.py in the build container on scripted random draws and
 *       comparing with this file (tests/test_reference_live_cpu.py, tests/ref_live.py); the vectors travel as
 *       tests/golden/ref_live_replay.npz / ref_live_trace.json (tests/golden/make_ref_live_golden.py) and are checked on
 *       the CPU (tests/test_oracle_cpu.py) and on the B200 (tests/test_gpu_reference_golden.py).  Bit-exact, with one
 *       stated exception: the stretch move's np.exp is libm's, this file (like the device) uses the deterministic hd_exp
 *       (<= 2 ulp apart); with that single substitution whole runs are bit-identical, without it decisions are identical
 *       and states agree to 1e-13.  The shear path evaluates np.dot the way numpy does in this image (FMA chain).
 *   (2) The entry points MC_run calls INTO hs_alkane (third-party Fortran, imported at MCNS.py:2, no version pinned,
 *       source not under /root/reference, gfortran / swig absent in this image) are restated from its published
 *       algorithm: hard spheres of diameter 1, triclinic minimum image through the reciprocal cell, rigid-chain
 *       translate / quaternion rotate / dihedral bond rotation, isotropic volume move with chains following by their
 *       first bead.  Every place where the real Fortran might differ is listed as assumption A1..A10 in DESIGN.md and
 *       flagged "[A#]" below; these remain unpinned until a box with hs_alkane exists.  They are constrained by the
 *       physical known-answer tests in tests/ (Carnahan-Starling equation of state, rigid bonds, volume preservation).
 *
 * Conventions: float64 throughout; H is a 3x3 cell, ROWS are the cell vectors
 * (MCNS.py:112-113, :50); positions are Cartesian row vectors, r = s.H; R is
 * [nchains][nbeads][3]; indices are 0-based here (the Python boundary is 1-based).
 */
#ifndef HANS_ORACLE_H
#define HANS_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { HO_VOL = 0, HO_TRANS = 1, HO_ROT = 2, HO_DIH = 3, HO_SHEAR = 4, HO_STRETCH = 5 };

/* reason codes of a decision */
enum {
    HO_ACCEPT = 0,
    HO_REJ_OVERLAP = 1,
    HO_REJ_CEILING = 2,
    HO_REJ_BOLTZ = 3,
    HO_REJ_ASPECT = 4,
    HO_REJ_ANGLE = 5,
    HO_REJ_INVALID = 6
};

typedef struct ho_cfg {
    int32_t nchains;
    int32_t nbeads;
    int32_t nexclude;   /* intra-chain pairs are tested iff |i-j| > nexclude  [A3], default 3 */
    int32_t allow_flip; /* dihedral move may reverse the chain first (MCNS.py:333 passes 1) */
    double dv_max, dr_max, dt_max, dh_max, dshear, dstretch;
    double min_ar;      /* MCNS.py:194 */
    double cos_min_ang; /* cos(min_ang*pi/180): arccos(c) < lim  <=>  c > cos(lim), MCNS.py:196 */
    double pressure;    /* MCNS.py:277, always 0 from mpihans.py */
    double move_cum[6]; /* cumsum(move_ratio)/sum(move_ratio), MCNS.py:294 */
    double bondlength;  /* MCNS.py:605 */
    double cos_bondangle, sin_bondangle; /* MCNS.py:606 (degrees at the boundary) */
} ho_cfg;

/* one explicit proposal */
typedef struct ho_proposal {
    int32_t type;   /* HO_VOL .. HO_STRETCH */
    int32_t ichain; /* drawn for every move (MCNS.py:308) */
    int32_t idx0;   /* dih: first rotated bead; shear: k; stretch: i */
    int32_t idx1;   /* dih: flip flag; stretch: j */
    double p[4];    /* vol: dV | trans: dr[3] | rot: quat (w,x,y,z) | dih: angle |
                       shear: rv1, rv2 | stretch: rv */
    double u_acc;   /* acceptance uniform (MCNS.py:351,372) */
    double reserved;
} ho_proposal;

typedef struct ho_decision {
    int32_t accepted;
    int32_t reason;
    double boltz;
} ho_decision;

/* modes of ho_apply_proposal */
enum {
    HO_MODE_FUSED = 0,  /* MC_run semantics: decide and revert on reject */
    HO_MODE_PROPOSE = 1 /* hs_alkane entry-point semantics: leave the trial state, return boltz */
};

void ho_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* include/hans_detmath.h exported for accuracy tests against libm */
double ho_det_exp(double x);
double ho_det_log(double x);
double ho_det_cbrt(double x);
void ho_det_sincos(double x, double *s, double *c);
double ho_det_rint(double x);

void ho_hinv(const double *H, double *Hi);
double ho_volume(const double *H);
double ho_min_aspect_ratio(const double *H);
double ho_max_abs_cos(const double *H);
int ho_aspect_fails(const double *H, double min_ar);
int ho_angle_fails(const double *H, double cos_min_ang);
double ho_pair_r2(const double *H, const double *Hi, const double *ri, const double *rj);
double ho_pair_r2_27(const double *H, const double *ri, const double *rj);
int ho_moves_per_sweep(int nbeads, int nchains);

int ho_check_overlap(const ho_cfg *c, const double *H, const double *R);       /* inter + intra */
int ho_check_overlap_inter(const ho_cfg *c, const double *H, const double *R); /* inter only */
int ho_chain_overlap(const ho_cfg *c, const double *H, const double *R, int ichain);

void ho_change_box(const ho_cfg *c, double *H, double *R, const double *dH);

void ho_gen_proposal(const ho_cfg *c, uint64_t seed, uint32_t gid, uint64_t ctr, ho_proposal *out);
void ho_apply_proposal(const ho_cfg *c, double *H, double *R, const ho_proposal *p,
                       double volume_limit, int mode, ho_decision *out);
void ho_replay(const ho_cfg *c, double *H, double *R, const ho_proposal *p, int nmoves,
               double volume_limit, int mode, ho_decision *out);

/* MC_run (MCNS.py:276-392), self-driving from Philox; returns the final volume */
double ho_mc_run(const ho_cfg *c, double *H, double *R, int sweeps, double volume_limit,
                 uint64_t seed, uint32_t gid, uint64_t *ctr, uint64_t attempted[6],
                 uint64_t accepted[6]);

/* the same for many walkers, one walker per OpenMP thread (CPU baseline "best case") */
void ho_mc_run_many(const ho_cfg *c, double *H, double *R, const int32_t *ids, int n, int sweeps,
                    const double *volume_limit, uint64_t seed, const uint32_t *gids,
                    uint64_t *ctrs, uint64_t *attempted, uint64_t *accepted, double *vols,
                    int nthreads);

/* chain growth [A9]; one attempt; returns ifail (0 ok) */
int ho_grow_chain(const ho_cfg *c, const double *H, double *R, int ichain, uint64_t seed,
                  uint32_t gid, uint32_t attempt);
int ho_grow_walker(const ho_cfg *c, const double *H, double *R, uint64_t seed, uint32_t gid,
                   int max_attempts);

#ifdef __cplusplus
}
#endif
#endif
