"""Minimal stand-in for the two ase pieces the reference uses on the hot path's boundary:
``ase.Atoms`` (MCNS.py:48-50, 566-572; NSio.py:121-122) and the extended-XYZ reader/writer
(``ase.io.write(..., append=True)`` NSio.py:124, ``ase.io.read`` mpihans.py:155, NSio.py:143).
ase is not installed in this image; if it is importable the NSio layer can use it instead."""
from __future__ import annotations

import re

import numpy as np


class Atoms:
    def __init__(self, symbols="", positions=None, pbc=True, cell=None):
        self.positions = np.array(positions, dtype=np.float64).reshape(-1, 3)
        c = np.array(cell if cell is not None else np.zeros((3, 3)), dtype=np.float64)
        self.cell = c * np.eye(3) if c.size == 3 else c.reshape(3, 3)
        self.pbc = np.array([bool(pbc)] * 3) if np.isscalar(pbc) else np.array(pbc, dtype=bool)
        m = re.fullmatch(r"([A-Z][a-z]?)(\d*)", symbols or "")
        sym = m.group(1) if m else "C"
        self.symbols = [sym] * len(self.positions)

    def __len__(self):
        return len(self.positions)

    def get_positions(self):
        return self.positions.copy()

    def get_number_of_atoms(self):
        return len(self.positions)

    def get_cell(self):
        return self.cell.copy()

    def get_volume(self):
        return float(abs(np.linalg.det(self.cell)))

    def wrap(self, eps=1e-7):
        """ase.Atoms.wrap (NSio.py:122): fractional coordinates into [-eps, 1-eps)."""
        frac = np.linalg.solve(self.cell.T, self.positions.T).T
        for k in range(3):
            if self.pbc[k]:
                frac[:, k] = (frac[:, k] + eps) % 1.0 - eps
        self.positions = frac @ self.cell


def write_extxyz(filename, images, append=False):
    """Frames in ase's extended-XYZ layout: N / Lattice="..." Properties=species:S:1:pos:R:3 pbc="T T T" /
    N lines of species + %16.8f coordinates."""
    if isinstance(images, Atoms):
        images = [images]
    with open(filename, "a" if append else "w") as f:
        for at in images:
            lat = " ".join(repr(float(x)) for x in at.cell.reshape(9))
            pbc = " ".join("T" if p else "F" for p in at.pbc)
            f.write(f"{len(at)}\n")
            f.write(f'Lattice="{lat}" Properties=species:S:1:pos:R:3 pbc="{pbc}"\n')
            for s, p in zip(at.symbols, at.positions):
                f.write("%-2s %16.8f %16.8f %16.8f\n" % (s, p[0], p[1], p[2]))


def read_extxyz(filename, index=None):
    """Read frames written by write_extxyz / ase.  index None -> last frame (ase.io.read default),
    a slice string like ':5' or a slice -> list of frames."""
    frames = []
    with open(filename) as f:
        while True:
            line = f.readline()
            if not line:
                break
            if not line.strip():
                continue
            n = int(line.split()[0])
            comment = f.readline()
            m = re.search(r'Lattice="([^"]*)"', comment)
            cell = np.array([float(x) for x in m.group(1).split()]).reshape(3, 3) if m else np.zeros((3, 3))
            m = re.search(r'pbc="([^"]*)"', comment)
            pbc = [t.upper().startswith("T") for t in m.group(1).split()] if m else [True] * 3
            sym, pos = [], []
            for _ in range(n):
                p = f.readline().split()
                sym.append(p[0])
                pos.append([float(p[1]), float(p[2]), float(p[3])])
            at = Atoms(f"{sym[0]}{n}" if sym else "", positions=pos, pbc=pbc, cell=cell)
            at.symbols = sym
            frames.append(at)
    if index is None:
        return frames[-1]
    if isinstance(index, str):
        parts = [int(p) if p else None for p in index.split(":")]
        index = slice(*parts) if len(parts) > 1 else parts[0]
    return frames[index]
