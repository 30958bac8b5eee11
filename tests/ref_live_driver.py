"""Runs the UNMODIFIED reference driver /root/reference/mpihans.py (and the reference's own NesSa/NSio.py) in this container
on one fake MPI rank (TEST INFRASTRUCTURE; never imported by hans_b200/).

Stand-ins put into sys.modules for the import:
  mpi4py.MPI   COMM_WORLD of size 1 (bcast / allreduce(MAXLOC) / Allreduce / Barrier / Wtime are identities)
  hs_alkane    tests/ref_live.OracleAlk driven by its own numpy generator (free-running: the moves' amplitudes follow the
               oracle's proposal laws A4-A8; nothing is scripted here -- this harness pins the FILE CONTRACT, not decisions)
  ase          an Atoms with wrap(), ase.io.write / read that RECORD what the reference hands them (and keep frames in memory)
  h5py         a File whose "w" mode records attrs / groups / datasets exactly as the reference creates them and, on close, writes
               them with hans_b200.h5min -- the package's own restart writer -- and whose "r" mode serves what hans_b200.h5min
               reads back: so the reference's reader (mpihans.py:77-89,166-181) is executed on files this package writes.
"""
from __future__ import annotations

import contextlib
import importlib
import io
import os
import sys
import types

import numpy as np

import ref_live
from oracle import oracle as ho

REFERENCE_ROOT = ref_live.REFERENCE_ROOT


# ---- hs_alkane: free-running -----------------------------------------------------------------------------------------
class FreeAlk(ref_live.OracleAlk):
    """OracleAlk whose random_uniform_random and move internals come from a numpy generator (oracle proposal laws)."""

    def __init__(self, seed=0):
        super().__init__()
        self.rng = np.random.default_rng(seed)

    def random_uniform_random(self):
        return float(self.rng.random())

    def _pop_internal(self, kind, w, ichain=None):
        p = np.zeros(1, dtype=ho.PROPOSAL_DTYPE)
        p["type"] = kind
        if ichain is not None:
            p["ichain"] = ichain
        u = self.rng.random(4)
        if kind == ho.VOL:
            p["p"][0, 0] = self.steps["dv_max"] * (2 * u[0] - 1)
        elif kind == ho.TRANS:
            p["p"][0, :3] = self.steps["dr_max"] * (2 * u[:3] - 1)
        elif kind == ho.ROT:
            ax = 2 * u[:3] - 1
            ax /= np.linalg.norm(ax)
            th = min(self.steps["dt_max"], np.pi) * (2 * u[3] - 1)
            p["p"][0] = [np.cos(th / 2), *(np.sin(th / 2) * ax)]
        else:
            nb = self.nbeads
            p["idx0"] = 3 + min(int(u[0] * (nb - 3)), nb - 4) if nb >= 4 else -1
            p["idx1"] = int(u[1] < 0.5)
            p["p"][0, 0] = min(self.steps["dh_max"], np.pi) * (2 * u[2] - 1)
        return p

    def alkane_bond_rotate(self, ichain, ibox, allow_flip):
        if self.nbeads < 4:
            return 0.0, 0, 0.0
        return super().alkane_bond_rotate(ichain, ibox, allow_flip)


# ---- ase ---------------------------------------------------------------------------------------------------------------
class RecAtoms(ref_live.StubAtoms):
    def wrap(self, eps=1e-7):
        frac = np.linalg.solve(self.cell.T, self.positions.T).T
        frac = (frac + eps) % 1.0 - eps
        self.positions = frac @ self.cell
        self.wrapped = True

    def get_number_of_atoms(self):
        return len(self.positions)


class AseIO(types.ModuleType):
    def __init__(self):
        super().__init__("ase.io")
        self.files = {}     # filename -> list of frames (cell, positions)
        self.calls = []     # ("write", filename, append, nframes) / ("read", filename, index)

    def write(self, filename, images, append=False, parallel=True, format=None):
        frames = images if isinstance(images, list) else [images]
        key = os.path.abspath(filename)
        if not append:
            self.files[key] = []
        self.files.setdefault(key, []).extend((np.array(a.cell), np.array(a.positions), getattr(a, "wrapped", False)) for a in frames)
        self.calls.append(("write", os.path.basename(filename), bool(append), len(frames)))

    def read(self, filename, index=None, parallel=True):
        key = os.path.abspath(filename)
        self.calls.append(("read", os.path.basename(filename), index))
        frames = [RecAtoms("C", positions=p, pbc=True, cell=c) for c, p, _ in self.files.get(key, [])]
        if index is None:
            if not frames:
                raise OSError(filename)
            return frames[-1]
        assert isinstance(index, str) and index.startswith(":")
        return frames[:int(index[1:])]


# ---- h5py on top of hans_b200.h5min ----------------------------------------------------------------------------------------
class _Attrs(dict):
    def create(self, k, v):
        self[k] = v


class _Dataset:
    def __init__(self, shape, dtype):
        assert dtype == "float64"
        self.a = np.zeros(shape, dtype=np.float64)

    def __setitem__(self, k, v):
        self.a[k] = v

    def __getitem__(self, k):
        return self.a[k]


class _Group(dict):
    def create_dataset(self, name, shape, dtype=None):
        self[name] = _Dataset(shape, dtype)
        return self[name]


class FakeH5File:
    log = []   # ("w" | "r", basename) of every open, for the tests

    def __init__(self, filename, mode="r"):
        from hans_b200 import h5min
        self.filename, self.mode = filename, mode
        self.attrs, self.groups = _Attrs(), {}
        FakeH5File.log.append((mode, os.path.basename(filename)))
        if mode == "r":
            attrs, groups = h5min.read(filename)   # the reference reads what THIS PACKAGE's writer wrote
            for k, v in attrs.items():
                if isinstance(v, np.ndarray) and v.dtype.kind in "US" and v.shape == ():
                    v = str(v)
                self.attrs[k] = v
            for g, d in groups.items():
                grp = _Group()
                for name, arr in d.items():
                    ds = _Dataset(arr.shape, "float64")
                    ds.a[...] = arr
                    grp[name] = ds
                self.groups[g] = grp

    def create_group(self, name):
        self.groups[name] = _Group()
        return self.groups[name]

    def __getitem__(self, name):
        return self.groups[name]

    def __iter__(self):
        return iter(self.groups)

    def close(self):
        if self.mode == "w":
            from hans_b200 import h5min
            h5min.write(self.filename, dict(self.attrs), {g: {n: d.a for n, d in grp.items()} for g, grp in self.groups.items()})


# ---- mpi4py ---------------------------------------------------------------------------------------------------------------
class _Comm:
    def Get_size(self): return 1
    def Get_rank(self): return 0
    def bcast(self, obj, root=0): return obj
    def allreduce(self, obj, op=None): return obj
    def Allreduce(self, src, dst, op=None): dst[...] = src
    def Barrier(self): pass
    def ssend(self, *a, **k): raise AssertionError("size 1: nothing is sent")
    def recv(self, *a, **k): raise AssertionError("size 1: nothing is received")


class LiveDriver:
    """The reference's mpihans / NesSa.NSio / NesSa.MCNS modules, imported unmodified with the stand-ins above."""

    def __init__(self, seed=0):
        if not ref_live.reference_available():
            raise FileNotFoundError(REFERENCE_ROOT)
        self.alk = FreeAlk(seed)
        self.aseio = AseIO()
        names = ("hs_alkane", "hs_alkane.alkane", "ase", "ase.io", "ase.visualize", "mpi4py", "h5py", "NesSa", "NesSa.MCNS",
                 "NesSa.NSio", "NesSa.ns_analyse", "NesSa.ns_analyse_main", "_hans_reference_mpihans")
        saved = {k: sys.modules.get(k) for k in names}
        saved_path = list(sys.path)
        try:
            for k in names:
                sys.modules.pop(k, None)
            hs = types.ModuleType("hs_alkane"); hs.alkane = self.alk
            ase = types.ModuleType("ase"); ase.Atoms = RecAtoms; ase.io = self.aseio
            ase.visualize = types.ModuleType("ase.visualize"); ase.visualize.view = lambda *a, **k: None
            mpi = types.ModuleType("mpi4py")
            import time as _t
            mpi.MPI = types.SimpleNamespace(COMM_WORLD=_Comm(), MAXLOC="MAXLOC", SUM="SUM", Wtime=_t.time)
            h5 = types.ModuleType("h5py"); h5.File = FakeH5File
            sys.modules.update({"hs_alkane": hs, "hs_alkane.alkane": self.alk, "ase": ase, "ase.io": self.aseio,
                                "ase.visualize": ase.visualize, "mpi4py": mpi, "h5py": h5})
            sys.path.insert(0, REFERENCE_ROOT)
            self.NS = importlib.import_module("NesSa.MCNS")
            self.NSio = importlib.import_module("NesSa.NSio")
            spec = importlib.util.spec_from_file_location("_hans_reference_mpihans", os.path.join(REFERENCE_ROOT, "mpihans.py"))
            self.mpihans = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(self.mpihans)
        finally:
            sys.path[:] = saved_path
            for k, v in saved.items():
                if v is None:
                    sys.modules.pop(k, None)
                else:
                    sys.modules[k] = v
        assert self.NS.alk is self.alk
        self.alk.reference = self.NS
        self.NS.io = self.aseio   # MCNS.py:6 `from ase import io` (NSio.write_to_extxyz calls NS.io.write)

    def parse(self, text):
        """NesSa.NSio.read_hans_file on stdin."""
        old = sys.stdin
        sys.stdin = io.StringIO(text)
        try:
            return self.NSio.read_hans_file()
        finally:
            sys.stdin = old

    def run(self, text, cwd):
        """`python mpihans.py < input` in directory cwd; returns (SimParams as main() left them, captured stdout)."""
        params = self.parse(text)
        buf = io.StringIO()
        here = os.getcwd()
        os.chdir(cwd)
        self.alk._reset()
        self.alk.cfg_kw = dict(min_ar=params["min_aspect_ratio"], min_ang=params["min_angle"])
        try:
            with contextlib.redirect_stdout(buf):
                self.mpihans.main(params)
        finally:
            os.chdir(here)
        return params, buf.getvalue()
