"""Harness that executes the UNMODIFIED reference module /root/reference/NesSa/MCNS.py in this
container (TEST INFRASTRUCTURE; never imported by hans_b200/).

The reference imports `hs_alkane.alkane`, `ase` and `mpi4py` at module level (MCNS.py:2-19); none of
them exists in this image.  `load_reference()` puts stand-ins into sys.modules and imports the file
as it lies on disk.  The stand-in for `hs_alkane.alkane` is `OracleAlk`: the 39 entry points of
SURVEY.md 8(b) on top of the CPU oracle's C functions (oracle/hans_oracle.c), with

  * the two random streams of the reference SCRIPTED: `alk.random_uniform_random()` (stream F,
    MCNS.py:148,222,223,308) and `np.random.random / normal` (stream P, MCNS.py:174,175,227,313,
    351,372) pop from queues the test fills, so a run is a pure function of (state, script);
  * the random internals of hs_alkane's own move routines (translate / rotate / dihedral / volume
    amplitudes -- unknowable without the Fortran) taken from a queue of
    explicit proposal records;
  * a call trace (entry point, arguments, results, Python-side writes through the chain views)
    that the GPU tests replay against hans_b200.alk (tests/golden/ref_live_*.json).

What this pins: everything MCNS.py itself computes -- move selection thresholds, draw order and
count, box_shear_step / box_stretch_step arithmetic (delta_H, rejection order), min_aspect_ratio,
min_angle, accept / revert / ceiling logic of MC_run, rate bookkeeping -- against the oracle's
restatement (ho_apply_proposal in FUSED mode) on identical draws.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types
from collections import deque

import numpy as np

from oracle import oracle as ho

REFERENCE_ROOT = "/root/reference"
REFERENCE_MCNS = os.path.join(REFERENCE_ROOT, "NesSa", "MCNS.py")


def reference_available() -> bool:
    return os.path.exists(REFERENCE_MCNS)


# ---------------------------------------------------------------------------------------------------
# stand-ins for the absent third-party modules
# ---------------------------------------------------------------------------------------------------
class StubAtoms:
    """The handful of ase.Atoms features MCNS.py touches (MCNS.py:48-50,564-577)."""

    def __init__(self, symbols=None, positions=None, pbc=None, cell=None):
        self.symbols = symbols
        self.positions = np.array(positions, dtype=np.float64)
        self.cell = np.array(cell, dtype=np.float64)
        self.pbc = pbc

    def get_positions(self):
        return self.positions.copy()

    def __len__(self):
        return len(self.positions)


class ScriptExhausted(AssertionError):
    pass


class ScriptedRandom:
    """`np.random` as MCNS.py sees it: random(), normal(scale=), randint(n) pop scripted values."""

    def __init__(self):
        self.uniform = deque()
        self.normals = deque()   # (expected scale, value to return)
        self.ints = deque()
        self.calls = 0

    def random(self):
        self.calls += 1
        if not self.uniform:
            raise ScriptExhausted("np.random.random() called more often than scripted")
        return self.uniform.popleft()

    def normal(self, loc=0.0, scale=1.0):
        self.calls += 1
        if not self.normals:
            raise ScriptExhausted("np.random.normal() called more often than scripted")
        want_scale, value = self.normals.popleft()
        assert loc == 0.0 and scale == want_scale, f"normal(scale={scale}) but the script expects scale {want_scale}"
        return value

    def randint(self, n):
        self.calls += 1
        if not self.ints:
            raise ScriptExhausted("np.random.randint() called more often than scripted")
        v = self.ints.popleft()
        assert 0 <= v < n
        return v

    def empty(self):
        return not (self.uniform or self.normals or self.ints)


class NumpyProxy:
    """numpy with `.random` replaced; everything else (dot, sqrt, exp, cbrt, arccos ...) is numpy's own."""

    def __init__(self, rnd: ScriptedRandom, overrides=None):
        self.random = rnd
        self._ov = dict(overrides or {})

    def __getattr__(self, name):
        if name in self._ov:
            return self._ov[name]
        return getattr(np, name)


class OracleAlk(types.ModuleType):
    """`hs_alkane.alkane` on the CPU oracle (see the module docstring)."""

    def __init__(self):
        super().__init__("hs_alkane.alkane")
        self.F = deque()          # scripted alk.random_uniform_random() values
        self.internal = deque()   # scripted proposal records of the hs_alkane move routines
        self.trace = []           # (name, args, result) of every call that touches a box
        self.tracing = False
        self._reset()

    def _reset(self):
        self.nbox = 1
        self.nchains = 1
        self.nbeads = 1
        self.bondlength, self.bondangle = 0.4, 109.7
        self.steps = dict(dv_max=2.0, dr_max=2.0, dt_max=0.43, dh_max=0.4)
        self.H = self.R = None
        self.saved = {}
        self.cfg_kw = {}
        self.grow_attempt = 0
        self.seed = 0
        self.pressure = 0.0
        self._last_state = None

    # -- helpers ------------------------------------------------------------------------------------
    def _cfg(self, pressure=0.0):
        kw = dict(self.cfg_kw)
        kw.update(self.steps)
        kw.setdefault("min_ar", 0.8)
        kw.setdefault("min_ang", 45.0)
        return ho.make_cfg(self.nchains_alloc, self.nbeads, [1, 1, 1, 1, 1, 1], pressure=pressure,
                           bondlength=self.bondlength, bondangle=self.bondangle, **kw)

    def _note_python_writes(self, w):
        """A caller may have written through a chain view / cell view since the last entry point
        returned (MCNS.py:379-380): record that as an explicit event so a replay can repeat it."""
        if not self.tracing:
            return
        if self._last_state is not None and self._last_state[0] == w:
            _, H0, R0 = self._last_state
            if not np.array_equal(H0, self.H[w]):
                self.trace.append(("py_write_cell", [w + 1], self.H[w].tolist()))
            if not np.array_equal(R0, self.R[w]):
                nb = self.nbeads
                for c in range(self.nchains_alloc):
                    a, b = R0[c * nb:(c + 1) * nb], self.R[w][c * nb:(c + 1) * nb]
                    if not np.array_equal(a, b):
                        self.trace.append(("py_write_chain", [c + 1, w + 1], b.tolist()))

    def _rec(self, name, args, result, w=None):
        if self.tracing:
            self.trace.append((name, list(args), result))
            if w is not None:
                self._last_state = (w, self.H[w].copy(), self.R[w].copy())

    def _pop_internal(self, kind, w, ichain=None):
        if not self.internal:
            raise ScriptExhausted(f"hs_alkane routine {kind} called more often than scripted")
        p = self.internal.popleft()
        assert int(p["type"]) == kind, f"scripted proposal has type {int(p['type'])}, the reference asked for {kind}"
        if ichain is not None:
            assert int(p["ichain"]) == ichain, "the reference moved another chain than the script"
        return np.array([p], dtype=ho.PROPOSAL_DTYPE)

    def _propose(self, w, prop, pressure=0.0):
        dec = ho.replay(self._cfg(pressure), self.H[w], self.R[w], prop, mode=ho.MODE_PROPOSE)[0]
        return dec

    # -- lifecycle (MCNS.py:595-606) -------------------------------------------------------------------
    def box_set_quiet(self, q): pass
    def box_set_num_boxes(self, n): self.nbox = int(n)
    def box_initialise(self): pass
    def box_set_pbc(self, f): assert int(f) == 1
    def box_set_isotropic(self, f): assert int(f) == 1
    def box_set_bypass_link_cells(self, f): pass
    def box_set_use_verlet_list(self, f): pass
    def alkane_set_nchains(self, n): self.nchains = int(n)
    def alkane_set_nbeads(self, n): self.nbeads = int(n)

    def alkane_initialise(self):
        self.nchains_alloc = self.nchains
        N = self.nchains * self.nbeads
        self.H = np.zeros((self.nbox, 3, 3))
        self.H[:] = np.eye(3)
        self.R = np.zeros((self.nbox, N, 3))

    def alkane_set_bondlength(self, x): self.bondlength = float(x)
    def alkane_set_bondangle(self, x): self.bondangle = float(x)
    def alkane_destroy(self): self.H = self.R = None
    def box_destroy(self): pass

    # -- getters / setters -----------------------------------------------------------------------------
    def alkane_get_nchains(self): return self.nchains
    def alkane_get_nbeads(self): return self.nbeads
    def alkane_get_dv_max(self): return self.steps["dv_max"]
    def alkane_get_dr_max(self): return self.steps["dr_max"]
    def alkane_get_dt_max(self): return self.steps["dt_max"]
    def alkane_get_dh_max(self): return self.steps["dh_max"]
    def alkane_set_dv_max(self, x): self.steps["dv_max"] = float(x)
    def alkane_set_dr_max(self, x): self.steps["dr_max"] = float(x)
    def alkane_set_dt_max(self, x): self.steps["dt_max"] = float(x)
    def alkane_set_dh_max(self, x): self.steps["dh_max"] = float(x)

    # -- cell / chains ---------------------------------------------------------------------------------
    def box_get_cell(self, ibox):
        assert type(ibox) is int, "SWIG rejects numpy integers (MCNS.py:46)"
        return self.H[ibox - 1]

    def box_set_cell(self, ibox, cell):
        w = int(ibox) - 1
        self.H[w][...] = np.asarray(cell, dtype=np.float64).reshape(3, 3)
        self._rec("box_set_cell", [ibox, self.H[w].tolist()], None, w)

    def box_compute_volume(self, ibox):
        w = int(ibox) - 1
        self._note_python_writes(w)
        v = ho.volume(self.H[w])
        self._rec("box_compute_volume", [ibox], v, w)
        return v

    def box_min_aspect_ratio(self, ibox):
        """The fork's native routine; MCNS.py:97-118 is its Python specification in the same file."""
        w = int(ibox) - 1
        v = float(self.reference.min_aspect_ratio(int(ibox)))
        self._rec("box_min_aspect_ratio", [ibox], v, w)
        return v

    def alkane_get_chain(self, ichain, ibox):
        assert type(ichain) is int and type(ibox) is int, "SWIG rejects numpy integers (MCNS.py:310)"
        c = ichain - 1
        return self.R[ibox - 1][c * self.nbeads:(c + 1) * self.nbeads]

    # -- RNG -------------------------------------------------------------------------------------------
    def random_uniform_random(self):
        if not self.F:
            raise ScriptExhausted("alk.random_uniform_random() called more often than scripted")
        return self.F.popleft()

    def random_set_random_seed(self, s): self.seed = int(s)

    # -- moves -----------------------------------------------------------------------------------------
    def alkane_translate_chain(self, ichain, ibox):
        w = int(ibox) - 1
        self._note_python_writes(w)
        prop = self._pop_internal(ho.TRANS, w, ichain - 1)
        b = float(self._propose(w, prop)["boltz"])
        self._rec("alkane_translate_chain", [ichain, ibox, prop["p"][0, :3].tolist()], b, w)
        return b

    def alkane_rotate_chain(self, ichain, ibox, bond):
        w = int(ibox) - 1
        self._note_python_writes(w)
        prop = self._pop_internal(ho.ROT, w, ichain - 1)
        b = float(self._propose(w, prop)["boltz"])
        quat = prop["p"][0].copy()
        self._rec("alkane_rotate_chain", [ichain, ibox, int(bond), quat.tolist()], b, w)
        return b, quat

    def alkane_bond_rotate(self, ichain, ibox, allow_flip):
        w = int(ibox) - 1
        self._note_python_writes(w)
        prop = self._pop_internal(ho.DIH, w, ichain - 1)
        b = float(self._propose(w, prop)["boltz"])
        self._rec("alkane_bond_rotate", [ichain, ibox, int(allow_flip), int(prop["idx0"][0]), int(prop["idx1"][0]),
                                         float(prop["p"][0, 0])], b, w)
        return b, int(prop["idx0"][0]) + 1, float(prop["p"][0, 0])

    def alkane_box_resize(self, pressure, ibox, reset):
        w = int(ibox) - 1
        self._note_python_writes(w)
        if int(reset):
            Hs, Rs = self.saved.pop(w)
            self.H[w][...] = Hs
            self.R[w][...] = Rs
            self._rec("alkane_box_resize", [float(pressure), ibox, 1], 1.0, w)
            return 1.0
        self.saved[w] = (self.H[w].copy(), self.R[w].copy())
        prop = self._pop_internal(ho.VOL, w)
        b = float(self._propose(w, prop, pressure=float(pressure))["boltz"])
        self._rec("alkane_box_resize", [float(pressure), ibox, 0, float(prop["p"][0, 0])], b, w)
        return b

    def alkane_change_box(self, ibox, delta_H):
        w = int(ibox) - 1
        self._note_python_writes(w)
        dH = np.ascontiguousarray(delta_H, dtype=np.float64).reshape(3, 3)
        ho.change_box(self._cfg(), self.H[w], self.R[w], dH)
        self._rec("alkane_change_box", [ibox, dH.tolist()], None, w)

    def alkane_check_chain_overlap(self, ibox):
        w = int(ibox) - 1
        self._note_python_writes(w)
        r = int(ho.check_overlap(self._cfg(), self.H[w], self.R[w]))
        self._rec("alkane_check_chain_overlap", [ibox], r, w)
        return r

    def alkane_grow_chain(self, ichain, ibox, new_conf):
        w = int(ibox) - 1
        self.grow_attempt += 1
        fail = ho.grow_chain(self._cfg(), self.H[w], self.R[w], ichain - 1, self.seed, w, self.grow_attempt)
        return (0.0, 1) if fail else (1.0, 0)


class LiveReference:
    """The reference module + its scripted environment."""

    def __init__(self, np_overrides=None):
        if not reference_available():
            raise FileNotFoundError(REFERENCE_MCNS)
        self.alk = OracleAlk()
        self.rnd = ScriptedRandom()
        saved = {k: sys.modules.get(k) for k in ("hs_alkane", "hs_alkane.alkane", "ase", "ase.io", "ase.visualize", "mpi4py")}
        try:
            hs = types.ModuleType("hs_alkane")
            hs.alkane = self.alk
            ase = types.ModuleType("ase")
            ase.Atoms = StubAtoms
            ase.io = types.ModuleType("ase.io")
            ase.visualize = types.ModuleType("ase.visualize")
            ase.visualize.view = lambda *a, **k: None
            mpi = types.ModuleType("mpi4py")
            mpi.MPI = types.SimpleNamespace(SUM="SUM")
            sys.modules.update({"hs_alkane": hs, "hs_alkane.alkane": self.alk, "ase": ase, "ase.io": ase.io,
                                "ase.visualize": ase.visualize, "mpi4py": mpi})
            spec = importlib.util.spec_from_file_location("_hans_reference_MCNS", REFERENCE_MCNS)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)   # the file as it lies under /root/reference, unmodified
        finally:
            for k, v in saved.items():
                if v is None:
                    sys.modules.pop(k, None)
                else:
                    sys.modules[k] = v
        assert mod.alk is self.alk
        mod.np = NumpyProxy(self.rnd, np_overrides)   # scripted stream P; all numerics stay numpy's
        self.mod = mod
        self.alk.reference = mod

    # -- set-up ------------------------------------------------------------------------------------
    def setup(self, nchains, nbeads, H, R, steps=None, bondlength=0.4, bondangle=109.7):
        """initialise_sim_cells (MCNS.py:582-606) through the reference's own function, then the state."""
        H = np.asarray(H, dtype=np.float64).reshape(-1, 3, 3)
        R = np.asarray(R, dtype=np.float64).reshape(H.shape[0], nchains * nbeads, 3)
        self.alk._reset()
        args = dict(nwalkers=H.shape[0], nchains=nchains, nbeads=nbeads, bondlength=bondlength, bondangle=bondangle)
        self.mod.initialise_sim_cells(args, 1)
        if steps:
            for k, v in steps.items():
                getattr(self.alk, "alkane_set_" + k)(v)
        self.alk.H[...] = H
        self.alk.R[...] = R
        return args

    def script_moves(self, cfg, proposals, bump_stretch=False):
        """Fill streams F, P and the internal-proposal queue so that the reference's MC_run makes
        exactly `proposals` (records as the oracle / GPU consume them), in order."""
        nch = cfg.nchains
        cum = [cfg.move_cum[k] for k in range(6)]
        for p in proposals:
            t = int(p["type"])
            self.alk.F.append((int(p["ichain"]) + 0.5) / nch)                         # MCNS.py:308
            lo = cum[t - 1] if t > 0 else 0.0
            self.rnd.uniform.append(0.5 * (lo + cum[t]))                              # MCNS.py:313 (xi)
            if t in (ho.VOL, ho.TRANS, ho.ROT, ho.DIH):
                self.alk.internal.append(p.copy())
                self.rnd.uniform.append(float(p["u_acc"]))                            # MCNS.py:351,372
            elif t == ho.SHEAR:
                self.alk.F.append((int(p["idx0"]) + 0.5) / 3.0)                       # MCNS.py:148
                self.rnd.normals.append((cfg.dshear, float(p["p"][0])))               # MCNS.py:174
                self.rnd.normals.append((cfg.dshear, float(p["p"][1])))               # MCNS.py:175
            else:
                i, j = int(p["idx0"]), int(p["idx1"])
                self.alk.F.append((i + 0.5) / 3.0)                                    # MCNS.py:222
                # MCNS.py:223-225: j drawn independently, bumped to (j+1)%3 when it equals i
                if bump_stretch and j == (i + 1) % 3:
                    self.alk.F.append((i + 0.5) / 3.0)
                else:
                    self.alk.F.append((j + 0.5) / 3.0)
                self.rnd.normals.append((cfg.dstretch, float(p["p"][0])))             # MCNS.py:227

    def drained(self) -> bool:
        return not self.alk.F and not self.alk.internal and self.rnd.empty()

    def mc_run(self, cfg, proposals, move_ratio, ibox=1, volume_limit=sys.float_info.max, min_ang=45.0, pressure=0,
               trace=False, bump_stretch=False):
        """MC_run (MCNS.py:276-392) of the unmodified reference on box `ibox` with scripted draws."""
        mps = ho.moves_per_sweep(cfg.nbeads, cfg.nchains)
        assert len(proposals) % mps == 0, "MC_run walks whole sweeps"
        self.script_moves(cfg, proposals, bump_stretch)
        # the reference's thresholds np.cumsum(move_ratio)/np.sum(move_ratio) (MCNS.py:294) are what cfg.move_cum holds
        ratio = np.asarray(move_ratio, dtype=np.float64)
        assert np.array_equal(np.cumsum(ratio) / np.sum(ratio), [cfg.move_cum[k] for k in range(6)])
        ns_data = dict(nbeads=cfg.nbeads, nchains=cfg.nchains)
        self.alk.cfg_kw = dict(min_ar=cfg.min_ar, min_ang=min_ang, dshear=cfg.dshear, dstretch=cfg.dstretch,
                               nexclude=cfg.nexclude, allow_flip=cfg.allow_flip)
        self.alk.tracing = trace
        self.alk.trace = []
        self.alk._last_state = None
        try:
            vol, rates = self.mod.MC_run(ns_data, len(proposals) // mps, ratio, ibox, volume_limit,
                                         dshear=cfg.dshear, dstretch=cfg.dstretch, min_ang=min_ang, min_ar=cfg.min_ar,
                                         pressure=pressure)
        finally:
            self.alk.tracing = False
        return vol, rates


_cached = None


def live() -> LiveReference:
    """One import of the reference per process."""
    global _cached
    if _cached is None:
        _cached = LiveReference()
    return _cached
