"""ctypes binding of the CPU oracle (oracle/hans_oracle.c).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never by anything under hans_b200/.
Parity: pinned to the reference's own Python executed live (tests/test_reference_live_cpu.py); unpinned for
hs_alkane's Fortran internals (see oracle/hans_oracle.h).
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

VOL, TRANS, ROT, DIH, SHEAR, STRETCH = range(6)
MODE_FUSED, MODE_PROPOSE = 0, 1


class Cfg(C.Structure):
    _fields_ = [
        ("nchains", C.c_int32), ("nbeads", C.c_int32), ("nexclude", C.c_int32), ("allow_flip", C.c_int32),
        ("dv_max", C.c_double), ("dr_max", C.c_double), ("dt_max", C.c_double), ("dh_max", C.c_double),
        ("dshear", C.c_double), ("dstretch", C.c_double),
        ("min_ar", C.c_double), ("cos_min_ang", C.c_double), ("pressure", C.c_double),
        ("move_cum", C.c_double * 6),
        ("bondlength", C.c_double), ("cos_bondangle", C.c_double), ("sin_bondangle", C.c_double),
    ]


PROPOSAL_DTYPE = np.dtype([
    ("type", "<i4"), ("ichain", "<i4"), ("idx0", "<i4"), ("idx1", "<i4"),
    ("p", "<f8", (4,)), ("u_acc", "<f8"), ("reserved", "<f8"),
], align=True)
DECISION_DTYPE = np.dtype([("accepted", "<i4"), ("reason", "<i4"), ("boltz", "<f8")], align=True)
assert PROPOSAL_DTYPE.itemsize == 64 and DECISION_DTYPE.itemsize == 16


def make_cfg(nchains, nbeads, move_ratio, dv_max=2.0, dr_max=2.0, dt_max=0.43, dh_max=0.4, dshear=1.0,
             dstretch=1.0, min_ar=0.8, min_ang=45.0, pressure=0.0, bondlength=0.4, bondangle=109.7,
             nexclude=3, allow_flip=1) -> Cfg:
    """Defaults: step sizes mpihans.py:118-144, shape limits NSio.py:55-58, bond NSio.py:51-54."""
    mr = np.asarray(move_ratio, dtype=np.float64)
    cum = np.cumsum(mr) / np.sum(mr)  # MCNS.py:294
    c = Cfg()
    c.nchains, c.nbeads, c.nexclude, c.allow_flip = int(nchains), int(nbeads), int(nexclude), int(allow_flip)
    c.dv_max, c.dr_max, c.dt_max, c.dh_max = float(dv_max), float(dr_max), float(dt_max), float(dh_max)
    c.dshear, c.dstretch = float(dshear), float(dstretch)
    c.min_ar = float(min_ar)
    c.cos_min_ang = math.cos(float(min_ang) * math.pi / 180.0)
    c.pressure = float(pressure)
    for k in range(6):
        c.move_cum[k] = float(cum[k])
    c.bondlength = float(bondlength)
    ang = float(bondangle) * math.pi / 180.0
    c.cos_bondangle, c.sin_bondangle = math.cos(ang), math.sin(ang)
    return c


def _cpu_has_v3() -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    fl = set(line.split(":")[1].split())
                    return {"fma", "avx2", "bmi2"} <= fl
    except OSError:
        pass
    return False


def build(force: bool = False) -> None:
    """Compile the oracle with oracle/Makefile (gcc)."""
    if force:
        subprocess.run(["make", "-C", _HERE, "clean"], check=True, capture_output=True)
    subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    name = "libhans_oracle_v3.so" if _cpu_has_v3() else "libhans_oracle.so"
    path = os.path.join(_HERE, name)
    if not os.path.exists(path):
        build()
    L = C.CDLL(path)
    dp = C.POINTER(C.c_double)
    cp = C.POINTER(Cfg)
    vp = C.c_void_p
    L.ho_philox4x32_10.argtypes = [vp, vp, vp]
    L.ho_hinv.argtypes = [vp, vp]
    L.ho_volume.argtypes = [vp]; L.ho_volume.restype = C.c_double
    L.ho_min_aspect_ratio.argtypes = [vp]; L.ho_min_aspect_ratio.restype = C.c_double
    L.ho_max_abs_cos.argtypes = [vp]; L.ho_max_abs_cos.restype = C.c_double
    L.ho_aspect_fails.argtypes = [vp, C.c_double]; L.ho_aspect_fails.restype = C.c_int
    L.ho_angle_fails.argtypes = [vp, C.c_double]; L.ho_angle_fails.restype = C.c_int
    L.ho_pair_r2.argtypes = [vp, vp, vp, vp]; L.ho_pair_r2.restype = C.c_double
    L.ho_pair_r2_27.argtypes = [vp, vp, vp]; L.ho_pair_r2_27.restype = C.c_double
    L.ho_moves_per_sweep.argtypes = [C.c_int, C.c_int]; L.ho_moves_per_sweep.restype = C.c_int
    for fn in (L.ho_check_overlap, L.ho_check_overlap_inter):
        fn.argtypes = [cp, vp, vp]; fn.restype = C.c_int
    L.ho_chain_overlap.argtypes = [cp, vp, vp, C.c_int]; L.ho_chain_overlap.restype = C.c_int
    L.ho_change_box.argtypes = [cp, vp, vp, vp]
    L.ho_gen_proposal.argtypes = [cp, C.c_uint64, C.c_uint32, C.c_uint64, vp]
    L.ho_apply_proposal.argtypes = [cp, vp, vp, vp, C.c_double, C.c_int, vp]
    L.ho_replay.argtypes = [cp, vp, vp, vp, C.c_int, C.c_double, C.c_int, vp]
    L.ho_mc_run.argtypes = [cp, vp, vp, C.c_int, C.c_double, C.c_uint64, C.c_uint32, vp, vp, vp]
    L.ho_mc_run.restype = C.c_double
    L.ho_mc_run_many.argtypes = [cp, vp, vp, vp, C.c_int, C.c_int, vp, C.c_uint64, vp, vp, vp, vp, vp, C.c_int]
    L.ho_grow_chain.argtypes = [cp, vp, vp, C.c_int, C.c_uint64, C.c_uint32, C.c_uint32]
    L.ho_grow_chain.restype = C.c_int
    L.ho_grow_walker.argtypes = [cp, vp, vp, C.c_uint64, C.c_uint32, C.c_int]
    L.ho_grow_walker.restype = C.c_int
    for fn in (L.ho_det_exp, L.ho_det_log, L.ho_det_cbrt, L.ho_det_rint):
        fn.argtypes = [C.c_double]; fn.restype = C.c_double
    L.ho_det_sincos.argtypes = [C.c_double, dp, dp]
    _lib = L
    return L


def det_exp(x): return lib().ho_det_exp(float(x))
def det_log(x): return lib().ho_det_log(float(x))
def det_cbrt(x): return lib().ho_det_cbrt(float(x))
def det_rint(x): return lib().ho_det_rint(float(x))


def det_sincos(x):
    s, c = C.c_double(), C.c_double()
    lib().ho_det_sincos(float(x), C.byref(s), C.byref(c))
    return s.value, c.value


def _p(a: np.ndarray):
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data


def philox(ctr, key) -> np.ndarray:
    c = np.asarray(ctr, dtype=np.uint32).copy()
    k = np.asarray(key, dtype=np.uint32).copy()
    o = np.zeros(4, dtype=np.uint32)
    lib().ho_philox4x32_10(_p(c), _p(k), _p(o))
    return o


def hinv(H):
    H = np.ascontiguousarray(H, dtype=np.float64)
    Hi = np.empty((3, 3))
    lib().ho_hinv(_p(H), _p(Hi))
    return Hi


def volume(H) -> float:
    H = np.ascontiguousarray(H, dtype=np.float64)
    return lib().ho_volume(_p(H))


def min_aspect_ratio(H) -> float:
    H = np.ascontiguousarray(H, dtype=np.float64)
    return lib().ho_min_aspect_ratio(_p(H))


def max_abs_cos(H) -> float:
    H = np.ascontiguousarray(H, dtype=np.float64)
    return lib().ho_max_abs_cos(_p(H))


def aspect_fails(H, min_ar) -> int:
    H = np.ascontiguousarray(H, dtype=np.float64)
    return lib().ho_aspect_fails(_p(H), float(min_ar))


def angle_fails(H, cos_min_ang) -> int:
    H = np.ascontiguousarray(H, dtype=np.float64)
    return lib().ho_angle_fails(_p(H), float(cos_min_ang))


def pair_r2(H, ri, rj) -> float:
    H = np.ascontiguousarray(H, dtype=np.float64)
    Hi = hinv(H)
    ri = np.ascontiguousarray(ri, dtype=np.float64)
    rj = np.ascontiguousarray(rj, dtype=np.float64)
    return lib().ho_pair_r2(_p(H), _p(Hi), _p(ri), _p(rj))


def pair_r2_27(H, ri, rj) -> float:
    H = np.ascontiguousarray(H, dtype=np.float64)
    ri = np.ascontiguousarray(ri, dtype=np.float64)
    rj = np.ascontiguousarray(rj, dtype=np.float64)
    return lib().ho_pair_r2_27(_p(H), _p(ri), _p(rj))


def moves_per_sweep(nbeads, nchains) -> int:
    return lib().ho_moves_per_sweep(int(nbeads), int(nchains))


def check_overlap(cfg, H, R, inter_only=False) -> int:
    H = np.ascontiguousarray(H, dtype=np.float64)
    R = np.ascontiguousarray(R, dtype=np.float64)
    fn = lib().ho_check_overlap_inter if inter_only else lib().ho_check_overlap
    return fn(C.byref(cfg), _p(H), _p(R))


def chain_overlap(cfg, H, R, ichain) -> int:
    H = np.ascontiguousarray(H, dtype=np.float64)
    R = np.ascontiguousarray(R, dtype=np.float64)
    return lib().ho_chain_overlap(C.byref(cfg), _p(H), _p(R), int(ichain))


def change_box(cfg, H, R, dH) -> None:
    """In place on H (3,3) and R (N,3)."""
    dH = np.ascontiguousarray(dH, dtype=np.float64)
    lib().ho_change_box(C.byref(cfg), _p(H), _p(R), _p(dH))


def gen_proposals(cfg, seed, gid, ctr0, n) -> np.ndarray:
    out = np.zeros(n, dtype=PROPOSAL_DTYPE)
    L = lib()
    for m in range(n):
        L.ho_gen_proposal(C.byref(cfg), int(seed), int(gid), int(ctr0 + m), out[m:m + 1].ctypes.data)
    return out


def replay(cfg, H, R, proposals, volume_limit=float(np.finfo(np.float64).max), mode=MODE_FUSED) -> np.ndarray:
    """Apply explicit proposals in place to one walker (H (3,3), R (N,3)); returns decisions."""
    proposals = np.ascontiguousarray(proposals)
    assert proposals.dtype == PROPOSAL_DTYPE
    out = np.zeros(len(proposals), dtype=DECISION_DTYPE)
    lib().ho_replay(C.byref(cfg), _p(H), _p(R), _p(proposals), len(proposals), float(volume_limit), int(mode), _p(out))
    return out


def mc_run(cfg, H, R, sweeps, volume_limit, seed, gid, ctr):
    """MC_run on one walker in place.  Returns (volume, attempted[6], accepted[6], new ctr)."""
    att = np.zeros(6, dtype=np.uint64)
    acc = np.zeros(6, dtype=np.uint64)
    c = np.array([ctr], dtype=np.uint64)
    v = lib().ho_mc_run(C.byref(cfg), _p(H), _p(R), int(sweeps), float(volume_limit), int(seed), int(gid),
                        _p(c), _p(att), _p(acc))
    return v, att, acc, int(c[0])


def mc_run_many(cfg, H, R, ids, sweeps, volume_limit, seed, gids, ctrs, nthreads=0):
    """MC_run on walkers `ids` of the pool (H (nw,3,3), R (nw,N,3), ctrs (nw,)) in place, one per
    OpenMP thread.  Returns (vols[n], attempted[n,6], accepted[n,6])."""
    ids = np.ascontiguousarray(ids, dtype=np.int32)
    n = len(ids)
    gids = np.ascontiguousarray(gids, dtype=np.uint32)
    lim = np.ascontiguousarray(np.broadcast_to(np.asarray(volume_limit, dtype=np.float64), (n,)))
    att = np.zeros((n, 6), dtype=np.uint64)
    acc = np.zeros((n, 6), dtype=np.uint64)
    vols = np.zeros(n)
    assert ctrs.dtype == np.uint64
    lib().ho_mc_run_many(C.byref(cfg), _p(H), _p(R), _p(ids), n, int(sweeps), _p(lim), int(seed), _p(gids),
                         _p(ctrs), _p(att), _p(acc), _p(vols), int(nthreads))
    return vols, att, acc


def grow_chain(cfg, H, R, ichain, seed, gid, attempt) -> int:
    H = np.ascontiguousarray(H, dtype=np.float64)
    return lib().ho_grow_chain(C.byref(cfg), _p(H), _p(R), int(ichain), int(seed), int(gid), int(attempt))


def initial_cell(nbeads, nchains, max_vol_per_atom=15.0) -> np.ndarray:
    """create_initial_configs, MCNS.py:628-631."""
    return 0.999 * np.eye(3) * np.cbrt(nbeads * nchains * max_vol_per_atom)


def grow_walkers(cfg, nwalkers, seed, gid0=0, max_attempts=100000):
    """create_initial_configs + populate_boxes (MCNS.py:628-644) for `nwalkers` walkers."""
    N = cfg.nchains * cfg.nbeads
    H = np.ascontiguousarray(np.broadcast_to(initial_cell(cfg.nbeads, cfg.nchains), (nwalkers, 3, 3))).copy()
    R = np.zeros((nwalkers, N, 3))
    for w in range(nwalkers):
        r = lib().ho_grow_walker(C.byref(cfg), _p(H[w]), _p(R[w]), int(seed), int(gid0 + w), int(max_attempts))
        if r < 0:
            raise RuntimeError("oracle: chain growth failed")
    return H, R
