"""`alk` -- drop-in for the module object ``hs_alkane.alkane`` that hans imports at
/root/reference/NesSa/MCNS.py:2 (``import hs_alkane.alkane as alk``).

Same names, same argument meaning, same return shapes as the entry points the reference calls
(call sites cited per function).  Conventions kept from hs_alkane: `ibox` and
`ichain` are 1-based; cells are 3x3 float64 with ROWS = cell vectors; `alkane_get_chain` returns a
WRITABLE VIEW that callers mutate in place (MCNS.py:379-380,524,577; mpihans.py:177-179).

How it maps onto the GPU: hs_alkane's module arrays live here as host arrays (the views alias
them); every entry point that computes ships the affected box to the device pool, runs the
corresponding libhans_b200 kernel through the C ABI and brings the box back.  That makes this
layer latency-bound by design -- it exists so that NesSa-style scripts run unchanged.  The fast
path is `hans_b200.MCNS.MC_run`, which replaces the whole Python move loop with one fused kernel
launch, and `hans_b200.ns.NestedSampler`, which keeps the walkers resident on the device.

Randomness: hs_alkane draws its proposals from its own generator (random_uniform_random); here
they come from a numpy Generator seeded by random_set_random_seed (default seed 0) -- only
distributional equivalence is possible.
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np
import torch

from . import _lib
from ._lib import MODE_PROPOSE, PROPOSAL_DTYPE
from .engine import WalkerPool

_PI = math.pi


class _State:
    def __init__(self):
        self.quiet = 1
        self.nbox = 1
        self.pbc = 1
        self.isotropic = 1
        self.bypass_link_cells = 1
        self.use_verlet = 0
        self.nchains = 1
        self.nchains_alloc = None
        self.nbeads = 1
        self.bondlength = 0.4      # NSio.py:51
        self.bondangle = 109.7     # NSio.py:53
        self.steps = dict(dv_max=2.0, dr_max=2.0, dt_max=0.43, dh_max=0.4)  # mpihans.py:118-133
        self.pool: Optional[WalkerPool] = None
        self.H = None   # host master copies (views handed to callers alias these)
        self.R = None
        self.saved = {}  # ibox -> (H, R) snapshot of alkane_box_resize(.., reset=0)
        self.rng = np.random.default_rng(0)
        self.grow_attempt = 0
        self.seed = 0
        self.script = None  # explicit proposals for the move entry points (replay of a recorded call sequence)


_S = _State()


def _pool() -> WalkerPool:
    if _S.pool is None:
        raise _lib.HansError("alkane_initialise() has not been called")
    return _S.pool


def _sync_cfg() -> None:
    c = _S.pool.cfg
    c.dv_max, c.dr_max, c.dt_max, c.dh_max = (_S.steps[k] for k in ("dv_max", "dr_max", "dt_max", "dh_max"))
    c.bondlength = float(_S.bondlength)
    ang = float(_S.bondangle) * _PI / 180.0
    c.cos_bondangle, c.sin_bondangle = math.cos(ang), math.sin(ang)


def _w(ibox) -> int:
    w = int(ibox) - 1
    if not 0 <= w < _S.nbox:
        raise IndexError(f"ibox {ibox} out of range 1..{_S.nbox}")
    return w


def _c(ichain) -> int:
    c = int(ichain) - 1
    if not 0 <= c < _S.nchains_alloc:
        raise IndexError(f"ichain {ichain} out of range 1..{_S.nchains_alloc}")
    return c


def _up(w: int) -> None:
    p = _pool()
    p.H[w].copy_(torch.from_numpy(_S.H[w]))
    p.R[w].copy_(torch.from_numpy(_S.R[w]))


def _down(w: int) -> None:
    p = _pool()
    _S.H[w][...] = p.H[w].cpu().numpy()
    _S.R[w][...] = p.R[w].cpu().numpy()


# ---- lifecycle (MCNS.py:595-606, mpihans.py:301-302) -----------------------------------------------
def box_set_quiet(q): _S.quiet = int(q)
def box_set_num_boxes(n): _S.nbox = int(n)
def box_initialise(): pass


def box_set_pbc(flag):
    if not int(flag):
        raise NotImplementedError("hans_b200 implements periodic boundaries only (MCNS.py:598 sets pbc = 1)")
    _S.pbc = 1


def box_set_isotropic(flag):
    if not int(flag):
        raise NotImplementedError("only isotropic volume moves are implemented (MCNS.py:602 sets isotropic = 1)")
    _S.isotropic = 1


def box_set_bypass_link_cells(flag): _S.bypass_link_cells = int(flag)   # neighbour search is internal
def box_set_use_verlet_list(flag): _S.use_verlet = int(flag)


def alkane_set_nchains(n):
    """Before alkane_initialise: number of chains per box.  Afterwards only the logical count
    changes (populate_boxes lowers it while growing, MCNS.py:639); growth tests a new chain only
    against chains of lower index, so the temporary value needs no further effect."""
    n = int(n)
    if _S.pool is not None and n > _S.nchains_alloc:
        raise _lib.HansError("cannot raise nchains above the allocated value after alkane_initialise")
    _S.nchains = n


def alkane_set_nbeads(n):
    if _S.pool is not None and int(n) != _S.nbeads:
        raise _lib.HansError("cannot change nbeads after alkane_initialise")
    _S.nbeads = int(n)


def alkane_initialise():
    """Allocate the boxes on the device (replaces hs_alkane's module arrays)."""
    _S.nchains_alloc = _S.nchains
    N = _S.nchains * _S.nbeads
    _S.pool = WalkerPool(_S.nbox, _S.nchains, _S.nbeads, [1, 1, 1, 1, 1, 1], seed=_S.seed,
                         bondlength=_S.bondlength, bondangle=_S.bondangle, **_S.steps)
    _S.H = np.zeros((_S.nbox, 3, 3))
    _S.H[:] = np.eye(3)
    _S.R = np.zeros((_S.nbox, N, 3))
    _S.pool.set_state(_S.H, _S.R)
    _S.saved = {}
    _S.grow_attempt = 0


def alkane_set_bondlength(x):
    _S.bondlength = float(x)
    if _S.pool is not None:
        _sync_cfg()


def alkane_set_bondangle(x):
    _S.bondangle = float(x)
    if _S.pool is not None:
        _sync_cfg()


def alkane_destroy():
    _S.pool = None
    _S.H = _S.R = None
    _S.saved = {}


def box_destroy(): pass


# ---- getters / setters (MCNS.py:266,512-513,689-706; NSio.py:77-80; mpihans.py:119-133) ------------
def alkane_get_nchains(): return int(_S.nchains)
def alkane_get_nbeads(): return int(_S.nbeads)
def alkane_get_dv_max(): return float(_S.steps["dv_max"])
def alkane_get_dr_max(): return float(_S.steps["dr_max"])
def alkane_get_dt_max(): return float(_S.steps["dt_max"])
def alkane_get_dh_max(): return float(_S.steps["dh_max"])


def _set_step(name, x):
    _S.steps[name] = float(x)
    if _S.pool is not None:
        _sync_cfg()


def alkane_set_dv_max(x): _set_step("dv_max", x)
def alkane_set_dr_max(x): _set_step("dr_max", x)
def alkane_set_dt_max(x): _set_step("dt_max", x)
def alkane_set_dh_max(x): _set_step("dh_max", x)


# ---- cell (MCNS.py:43,105,194,519,570,631; mpihans.py:174,183) -------------------------------------
def box_get_cell(ibox) -> np.ndarray:
    _pool()
    return _S.H[_w(ibox)]


def box_set_cell(ibox, cell) -> None:
    _pool()
    _S.H[_w(ibox)][...] = np.asarray(cell, dtype=np.float64).reshape(3, 3)


def _metrics(w: int):
    p = _pool()
    _up(w)
    f64 = dict(dtype=torch.float64, device=p.device)
    out = torch.empty(3, **f64)
    _lib.check(_lib.lib.hans_cell_metrics(p.H[w].data_ptr(), 1, out[0:1].data_ptr(), out[1:2].data_ptr(),
                                          out[2:3].data_ptr(), torch.cuda.current_stream().cuda_stream))
    return out.cpu().numpy()


def box_compute_volume(ibox) -> float:
    return float(_metrics(_w(ibox))[0])


def box_min_aspect_ratio(ibox) -> float:
    return float(_metrics(_w(ibox))[1])


# ---- chains (MCNS.py:46,310,379-380,521-524,575-577; mpihans.py:177-179; NSio.py:94) ---------------
def alkane_get_chain(ichain, ibox) -> np.ndarray:
    """(nbeads,3) float64 WRITABLE VIEW of the chain's beads."""
    _pool()
    c = _c(ichain)
    return _S.R[_w(ibox)][c * _S.nbeads:(c + 1) * _S.nbeads]


# ---- RNG (MCNS.py:148,222,223,308) ------------------------------------------------------------------
def random_uniform_random() -> float:
    return float(_S.rng.random())


def random_set_random_seed(seed) -> None:
    _S.seed = int(seed)
    _S.rng = np.random.default_rng(int(seed))
    _S.grow_attempt = 0
    if _S.pool is not None:
        _S.pool.seed = int(seed)


# ---- moves (MCNS.py:186,238,318-333,357,368,642) ------------------------------------------------------
def set_proposal_script(records) -> None:
    """Parity hook (not part of hs_alkane): the next calls of alkane_translate_chain / alkane_rotate_chain /
    alkane_bond_rotate / alkane_box_resize(.., 0) take their random amplitudes from `records` (hans_proposal
    records, in call order) instead of the generator -- hs_alkane draws them inside the Fortran, so a recorded
    call sequence of the reference can only be replayed with the amplitudes made explicit.  None = off."""
    from collections import deque
    _S.script = None if records is None else deque(np.asarray(r, dtype=PROPOSAL_DTYPE).reshape(1).copy() for r in records)


def _scripted(kind: int, c=None):
    if _S.script is None:
        return None
    if not _S.script:
        raise _lib.HansError("proposal script exhausted")
    prop = _S.script.popleft()
    if int(prop["type"][0]) != kind or (c is not None and int(prop["ichain"][0]) != c):
        raise _lib.HansError("proposal script does not match the call (type %d / chain %s)" % (kind, c))
    return prop


def _propose(w: int, prop: np.ndarray, pressure: float = 0.0):
    """Run ONE explicit proposal on box w in hs_alkane entry-point semantics (trial state is left
    in place, the caller reverts): host box -> device, hans_replay(MODE_PROPOSE), device -> host."""
    p = _pool()
    _up(w)
    p.cfg.pressure = float(pressure)
    dec = p.replay([w], prop.reshape(1, 1), mode=MODE_PROPOSE)[0, 0]
    _down(w)
    return dec


def alkane_translate_chain(ichain, ibox) -> float:
    """Rigid shift of one chain by dr_max*(2U-1) per component; returns 0.0 on overlap else 1.0."""
    w, c = _w(ibox), _c(ichain)
    prop = _scripted(_lib.TRANS, c)
    if prop is None:
        prop = np.zeros(1, dtype=PROPOSAL_DTYPE)
        prop["type"], prop["ichain"] = _lib.TRANS, c
        prop["p"][0, :3] = _S.steps["dr_max"] * (2.0 * _S.rng.random(3) - 1.0)
    return float(_propose(w, prop)["boltz"])


def alkane_rotate_chain(ichain, ibox, bond=0):
    """Random rotation about the chain centroid by an angle in +-dt_max; bond != 0 rotates about the
    first bond vector instead of a random axis.  Returns (boltz, quaternion[4])."""
    w, c = _w(ibox), _c(ichain)
    nb = _S.nbeads
    prop = _scripted(_lib.ROT, c)
    if prop is not None:
        return float(_propose(w, prop)["boltz"]), prop["p"][0].copy()
    if int(bond) and nb >= 2:
        x = _S.R[w][c * nb:(c + 1) * nb]
        axis = x[1] - x[0]
    else:
        axis = 2.0 * _S.rng.random(3) - 1.0
    n = math.sqrt(float(axis @ axis))
    axis = axis / n if n > 0 else np.array([1.0, 0.0, 0.0])
    theta = min(_S.steps["dt_max"], _PI) * (2.0 * _S.rng.random() - 1.0)
    quat = np.array([math.cos(0.5 * theta), *(math.sin(0.5 * theta) * axis)])
    prop = np.zeros(1, dtype=PROPOSAL_DTYPE)
    prop["type"], prop["ichain"] = _lib.ROT, c
    prop["p"][0] = quat
    return float(_propose(w, prop)["boltz"]), quat


def alkane_bond_rotate(ichain, ibox, allow_flip=1):
    """Dihedral move: the tail beyond a random internal bond is rotated about that bond by an angle
    in +-dh_max (optionally after reversing the chain).  Returns (boltz, first rotated bead
    (1-based), angle)."""
    w, c = _w(ibox), _c(ichain)
    nb = _S.nbeads
    prop = _scripted(_lib.DIH, c)
    if prop is not None:
        return float(_propose(w, prop)["boltz"]), int(prop["idx0"][0]) + 1, float(prop["p"][0, 0])
    prop = np.zeros(1, dtype=PROPOSAL_DTYPE)
    prop["type"], prop["ichain"] = _lib.DIH, c
    if nb < 4:
        prop["idx0"] = -1
        return 0.0, 0, 0.0
    ia = 3 + min(int(_S.rng.random() * (nb - 3)), nb - 4)
    flip = int(bool(allow_flip) and _S.rng.random() < 0.5)
    angle = min(_S.steps["dh_max"], _PI) * (2.0 * _S.rng.random() - 1.0)
    prop["idx0"], prop["idx1"] = ia, flip
    prop["p"][0, 0] = angle
    return float(_propose(w, prop)["boltz"]), ia + 1, angle


def alkane_box_resize(pressure, ibox, reset=0) -> float:
    """reset = 0: save the box, propose an isotropic volume change dV = dv_max*(2U-1) with the chains
    following rigidly; returns 0.0 on overlap else exp(-P dV + nchains ln(V'/V)) (MCNS.py:266).
    reset = 1: restore the saved cell and chains (MCNS.py:357)."""
    w = _w(ibox)
    if int(reset):
        if w in _S.saved:
            H, R = _S.saved.pop(w)
            _S.H[w][...] = H
            _S.R[w][...] = R
        return 1.0
    _S.saved[w] = (_S.H[w].copy(), _S.R[w].copy())
    prop = _scripted(_lib.VOL)
    if prop is None:
        prop = np.zeros(1, dtype=PROPOSAL_DTYPE)
        prop["type"] = _lib.VOL
        prop["p"][0, 0] = _S.steps["dv_max"] * (2.0 * _S.rng.random() - 1.0)
    prop["u_acc"] = 0.0
    return float(_propose(w, prop, pressure=float(pressure))["boltz"])


def alkane_change_box(ibox, delta_H) -> None:
    """H += delta_H; chains follow the cell rigidly (MCNS.py:186,238,368)."""
    w = _w(ibox)
    p = _pool()
    _up(w)
    p.change_box([w], np.asarray(delta_H, dtype=np.float64).reshape(1, 3, 3))
    _down(w)


def _shape_step(ibox, prop: np.ndarray, min_ar: float, min_ang_deg: float):
    """box_shear_step / box_stretch_step (MCNS.py:135-252) for one explicit proposal: the box is left at the trial
    cell (the caller reverts, MCNS.py:366-368).  Returns (boltz 0/1, delta_H (3,3)) -- two launches:
    hans_shape_delta for the cell change the reference's Python builds, hans_replay(MODE_PROPOSE) for the move."""
    w = _w(ibox)
    p = _pool()
    _up(w)
    f64 = dict(dtype=torch.float64, device=p.device)
    d_ids = torch.tensor([w], dtype=torch.int32, device=p.device)
    d_prop = torch.as_tensor(prop.view(np.uint8).reshape(-1)).to(p.device)
    d_dH = torch.empty(9, **f64)
    _lib.check(_lib.lib.hans_shape_delta(p.H.data_ptr(), d_ids.data_ptr(), 1, d_prop.data_ptr(), d_dH.data_ptr(),
                                         torch.cuda.current_stream().cuda_stream))
    p.cfg.min_ar = float(min_ar)
    p.cfg.cos_min_ang = math.cos(float(min_ang_deg) * _PI / 180.0)
    dec = p.replay([w], prop.reshape(1, 1), mode=MODE_PROPOSE)[0, 0]
    _down(w)
    return int(dec["boltz"] != 0.0), d_dH.cpu().numpy().reshape(3, 3)


def alkane_box_scale(ibox, sa, sb, sc) -> None:
    """Scale the three cell vectors by (sa, sb, sc) (only called from dead code, MCNS.py:260)."""
    w = _w(ibox)
    s = np.array([float(sa), float(sb), float(sc)])[:, None]
    alkane_change_box(ibox, _S.H[w] * s - _S.H[w])


def alkane_grow_chain(ichain, ibox, new_conf=1):
    """One growth attempt for chain `ichain` against the chains of lower index (MCNS.py:642).
    Returns (rb_factor, ifail): (1.0, 0) if the chain was placed, (0.0, 1) if it overlapped."""
    w, c = _w(ibox), _c(ichain)
    p = _pool()
    _up(w)
    st = torch.zeros(1, dtype=torch.int32, device=p.device)
    _S.grow_attempt += 1
    _lib.check(_lib.lib.hans_grow_chain(p.cfg, p.H.data_ptr(), p.R.data_ptr(), w, c, _S.seed, p.gid_base + w,
                                        _S.grow_attempt, st.data_ptr(), torch.cuda.current_stream().cuda_stream))
    ok = int(st.item()) > 0
    if ok:
        _down(w)
    return (1.0, 0) if ok else (0.0, 1)


# ---- predicate (MCNS.py:198,247,543) ------------------------------------------------------------------
def alkane_check_chain_overlap(ibox) -> int:
    w = _w(ibox)
    p = _pool()
    _up(w)
    return int(p.check_overlap(ids=[w]).item())


# ---- batch helpers used by hans_b200.MCNS (not part of hs_alkane) -------------------------------------
def _upload_all() -> None:
    p = _pool()
    p.H.copy_(torch.from_numpy(_S.H))
    p.R.copy_(torch.from_numpy(_S.R))


def _download_all() -> None:
    p = _pool()
    _S.H[...] = p.H.cpu().numpy()
    _S.R[...] = p.R.cpu().numpy()
