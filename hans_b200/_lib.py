"""ctypes binding of libhans_b200.so (C ABI declared in include/hans_b200.h).

The library is the product: there is no Python / CPU fallback.  If the shared object has not been
built (``python -c "import __graft_entry__ as g; g.build()"``) importing this module raises.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# HANS_B200_LIB: another build of the same library (A/B measurements of compile-time variants)
LIB_PATH = os.environ.get("HANS_B200_LIB") or os.path.join(_HERE, "libhans_b200.so")

VOL, TRANS, ROT, DIH, SHEAR, STRETCH = range(6)
MODE_FUSED, MODE_PROPOSE = 0, 1
ACCEPT, REJ_OVERLAP, REJ_CEILING, REJ_BOLTZ, REJ_ASPECT, REJ_ANGLE, REJ_INVALID = range(7)
DBL_MAX = float(np.finfo(np.float64).max)


class HansError(RuntimeError):
    pass


class Cfg(C.Structure):
    """struct hans_cfg (include/hans_b200.h)."""
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

class NsArgs(C.Structure):
    """struct hans_ns_args (include/hans_b200.h)."""
    _fields_ = [
        ("N", C.c_int32), ("nlocal", C.c_int32), ("rank", C.c_int32), ("world", C.c_int32),
        ("nw_per_vrank", C.c_int32), ("traj_interval", C.c_int32), ("traj_slots", C.c_int32), ("log_cap", C.c_int32),
        ("seed", C.c_uint64),
        ("d_H", C.c_void_p), ("d_R", C.c_void_p), ("d_vol", C.c_void_p),
        ("d_iter", C.c_void_p), ("d_rec", C.c_void_p), ("d_gathered", C.c_void_p), ("d_limit", C.c_void_p),
        ("d_log", C.c_void_p), ("d_ids", C.c_void_p), ("d_traj", C.c_void_p), ("d_traj_iter", C.c_void_p),
    ]


EXPORTS = [
    "hans_abi_version", "hans_last_error", "hans_device_count", "hans_walker_smem_bytes",
    "hans_moves_per_sweep", "hans_default_threads_per_walker", "hans_default_window", "hans_default_team", "hans_mc_run_batch",
    "hans_mc_run_batch_ws", "hans_mc_run_launches", "hans_gen_proposals",
    "hans_replay",
    "hans_check_overlap", "hans_cell_metrics", "hans_change_box", "hans_shape_delta", "hans_grow", "hans_grow_chain",
    "hans_argmax", "hans_clone", "hans_ns_pack", "hans_ns_resolve", "hans_ns_peer_bytes", "hans_peer_alloc",
    "hans_peer_open", "hans_peer_close", "hans_peer_free", "hans_ns_exchange_peer", "hans_ns_peer_timeout", "hans_order_params", "hans_diag_counters",
    "hans_ffma_peak",
]


def make_cfg(nchains, nbeads, move_ratio, dv_max=2.0, dr_max=2.0, dt_max=0.43, dh_max=0.4, dshear=1.0,
             dstretch=1.0, min_ar=0.8, min_ang=45.0, pressure=0.0, bondlength=0.4, bondangle=109.7,
             nexclude=3, allow_flip=1) -> Cfg:
    """Defaults follow the reference: step sizes mpihans.py:118-144, shape limits NSio.py:55-58,
    bond geometry NSio.py:51-54; move_cum = cumsum(move_ratio)/sum (MCNS.py:294)."""
    mr = np.asarray(move_ratio, dtype=np.float64)
    if mr.shape != (6,):
        raise Exception("Move ratio array should be of length 6.")  # MCNS.py:614-615
    cum = np.cumsum(mr) / np.sum(mr)
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


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` from the repo root. "
            "hans_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, cp = C.c_void_p, C.POINTER(Cfg)
    i32, u32, u64, f64 = C.c_int, C.c_uint32, C.c_uint64, C.c_double
    L.hans_abi_version.restype = C.c_int
    L.hans_last_error.restype = C.c_char_p
    L.hans_device_count.restype = C.c_int
    L.hans_walker_smem_bytes.argtypes = [cp]; L.hans_walker_smem_bytes.restype = C.c_int64
    L.hans_moves_per_sweep.argtypes = [i32, i32]; L.hans_moves_per_sweep.restype = C.c_int
    L.hans_default_threads_per_walker.argtypes = [cp]; L.hans_default_threads_per_walker.restype = C.c_int
    L.hans_default_window.argtypes = [cp]; L.hans_default_window.restype = C.c_int
    L.hans_default_team.argtypes = [cp]; L.hans_default_team.restype = C.c_int
    L.hans_mc_run_launches.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint64]; L.hans_mc_run_launches.restype = C.c_int
    L.hans_mc_run_batch.argtypes = [cp, vp, vp, vp, vp, i32, i32, vp, f64, u64, u32, vp, vp, vp, vp, i32, vp]
    L.hans_mc_run_batch_ws.argtypes = [cp, vp, vp, vp, vp, i32, i32, vp, f64, u64, u32, vp, vp, vp, vp, i32, vp, u64, vp]
    L.hans_gen_proposals.argtypes = [cp, vp, vp, i32, i32, u64, u32, vp, vp]
    L.hans_replay.argtypes = [cp, vp, vp, vp, i32, vp, i32, f64, i32, vp, i32, vp]
    L.hans_check_overlap.argtypes = [cp, vp, vp, vp, i32, i32, vp, i32, vp]
    L.hans_cell_metrics.argtypes = [vp, i32, vp, vp, vp, vp]
    L.hans_change_box.argtypes = [cp, vp, vp, vp, i32, vp, vp]
    L.hans_shape_delta.argtypes = [vp, vp, i32, vp, vp, vp]
    L.hans_grow.argtypes = [cp, vp, vp, i32, u64, u32, i32, vp, vp]
    L.hans_grow_chain.argtypes = [cp, vp, vp, i32, i32, u64, u32, u32, vp, vp]
    L.hans_argmax.argtypes = [vp, i32, vp, vp]
    L.hans_clone.argtypes = [cp, vp, vp, vp, i32, vp, vp, vp, vp, i32, vp, vp, vp]
    L.hans_ns_pack.argtypes = [C.POINTER(NsArgs), vp]
    L.hans_ns_resolve.argtypes = [C.POINTER(NsArgs), vp]
    L.hans_order_params.argtypes = [cp, vp, vp, i32, vp, vp, vp]
    L.hans_diag_counters.argtypes = [vp, i32]
    L.hans_ns_peer_bytes.argtypes = [i32, i32]; L.hans_ns_peer_bytes.restype = C.c_uint64
    L.hans_peer_alloc.argtypes = [u64, C.POINTER(vp), C.c_char_p]
    L.hans_peer_open.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.hans_peer_close.argtypes = [vp]
    L.hans_peer_free.argtypes = [vp]
    L.hans_ns_exchange_peer.argtypes = [C.POINTER(NsArgs), C.POINTER(vp), vp]
    L.hans_ns_peer_timeout.argtypes = [f64]
    L.hans_ffma_peak.argtypes = [i32, C.POINTER(f64), vp]
    for name in EXPORTS:
        fn = getattr(L, name)
        if name not in ("hans_last_error", "hans_walker_smem_bytes"):
            fn.restype = C.c_int
    if L.hans_abi_version() != 1:
        raise ImportError("libhans_b200.so ABI version mismatch; rebuild")
    return L


lib = _load()


def check(rc: int) -> None:
    if rc != 0:
        raise HansError(f"libhans_b200 error {rc}: {lib.hans_last_error().decode()}")


def diag_counters(reset: bool = False):
    """(float64 chain re-tests, float64 all-pairs passes) since the last reset."""
    out = (C.c_uint64 * 4)()
    check(lib.hans_diag_counters(out, int(reset)))
    return [int(x) for x in out]


def moves_per_sweep(nbeads: int, nchains: int) -> int:
    return lib.hans_moves_per_sweep(int(nbeads), int(nchains))
