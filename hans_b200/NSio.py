"""Mirror of the reference's ``NesSa.NSio``: stdin `key = value` input, restart file, extxyz
trajectory (all file:line citations into /root/reference/NesSa/NSio.py unless noted).

restart.hdf5: schema of NSio.py:69-111 -- root attributes = every parameter + prev_iters + step
sizes, one group per walker `walker_{rank}_{i:04d}` with `coordinates (N,3) f8` and
`unitcell (3,3) f8`.  Written with h5py when it is importable; h5py is NOT installed in this image,
so the file is then written by `hans_b200/h5min.py`, a dependency-free writer of the HDF5 subset the
schema needs (superblock 0, old-style groups, contiguous float64 datasets, scalar / 1-d int64,
float64 and fixed-length string attributes).  `open_restart` reads h5py-written files through h5py,
h5min-written files through h5min, and the `.npz` container earlier versions of this package wrote.
"""
from __future__ import annotations

import os
import sys

import numpy as np

from . import h5min
from .atoms import Atoms, read_extxyz, write_extxyz

try:  # pragma: no cover - not available in the build image
    import h5py  # type: ignore
    HAVE_H5PY = True
except ImportError:
    h5py = None
    HAVE_H5PY = False

FLOAT_KEYS = ["bondlength", "bondangle", "min_angle", "min_aspect_ratio", "pressure", "upper_bound", "lower_bound", "time"]
INT_KEYS = ["nchains", "nbeads", "nwalkers", "walklength", "initial_walk", "analyse", "equil_iter", "main_iter", "index"]
BOOL_KEYS = ["profile"]
# keys of this implementation (not in the reference): all optional
EXTRA_INT_KEYS = ["seed", "ranks_per_gpu", "restart_interval", "traj_interval", "threads_per_walker"]


def parse_hans_text(f: str) -> dict:
    """read_hans_file's parsing (NSio.py:29-62) on a string."""
    data = {}
    if f == "":
        print("Error: No input file specified")
        sys.exit(1)
    for line in f.split("\n"):
        if not line.startswith('#') and line.strip() != '':
            key, value = line.split("=", 1)
            data[key.strip()] = value.strip()
    for key in FLOAT_KEYS:
        if key in data:
            data[key] = float(data[key])
    for key in INT_KEYS + EXTRA_INT_KEYS:
        if key in data:
            data[key] = int(data[key])
    for key in BOOL_KEYS:
        if key in data:
            data[key] = bool(int(data[key]))
    data.setdefault("initial_walk", 1000)
    data.setdefault("analyse", 0)
    data.setdefault("bondlength", 0.4)
    data.setdefault("bondangle", 109.7)
    data.setdefault("min_angle", 45.0)
    data.setdefault("min_aspect_ratio", 0.8)
    data.setdefault("restart_file", "restart.hdf5")
    return data


def read_hans_file(filename=None) -> dict:
    """NSio.py:19-67: parse the input from stdin (default) or a file."""
    if filename is None:
        f = sys.stdin.read()
    else:
        with open(filename, "tr") as inp_file:
            f = inp_file.read()
    return parse_hans_text(f)


# ---- restart container ------------------------------------------------------------------------------
def _npz_name(filename: str) -> str:
    return filename if filename.endswith(".npz") else filename + ".npz"


def restart_exists(filename: str) -> bool:
    return os.path.exists(filename) or os.path.exists(_npz_name(filename))


def restart_paths(filename: str):
    """The on-disk names a restart called `filename` may occupy."""
    return [filename, _npz_name(filename)]


def _datasets(entry) -> dict:
    """(coordinates, unitcell[, philox_ctr]) -> the datasets of one walker group.  `philox_ctr` (one float64: the
    number of proposals the walker has consumed, exact below 2^53) is an addition of this implementation so that a
    resumed run continues the random streams instead of replaying them; the reference's reader ignores it."""
    d = {"coordinates": np.asarray(entry[0], dtype=np.float64).reshape(-1, 3),
         "unitcell": np.asarray(entry[1], dtype=np.float64).reshape(3, 3)}
    if len(entry) > 2 and entry[2] is not None:
        d["philox_ctr"] = np.asarray([float(entry[2])], dtype=np.float64)
    return d


def save_restart(filename: str, attrs: dict, walkers: dict) -> str:
    """walkers: {groupname: (coordinates (N,3), unitcell (3,3)[, philox counter])}.  Returns the path written."""
    if HAVE_H5PY:  # pragma: no cover
        with h5py.File(filename, "w") as f:
            for k, v in attrs.items():
                f.attrs.create(k, v)
            for g, entry in walkers.items():
                grp = f.create_group(g)
                for name, arr in _datasets(entry).items():
                    grp.create_dataset(name, data=arr)
        return filename
    tmp = filename + ".tmp"
    h5min.write(tmp, attrs, {g: _datasets(entry) for g, entry in walkers.items()})
    os.replace(tmp, filename)
    return filename


def open_restart(filename: str):
    """Returns (attrs, walkers) as written by save_restart (either container)."""
    if HAVE_H5PY and os.path.exists(filename):  # pragma: no cover
        attrs, walkers = {}, {}
        with h5py.File(filename, "r") as f:
            for k in f.attrs:
                attrs[k] = f.attrs[k]
            for g in f:
                ctr = float(f[g]["philox_ctr"][0]) if "philox_ctr" in f[g] else None
                walkers[g] = (f[g]["coordinates"][:], f[g]["unitcell"][:], ctr)
        return attrs, walkers
    if os.path.exists(filename) and h5min.is_hdf5(filename):
        try:
            attrs, groups = h5min.read(filename)
        except Exception as e:
            raise ValueError(f"{filename} is an HDF5 file this package's minimal reader cannot parse (written by libhdf5 "
                             f"with features outside the restart schema's subset?) and h5py is not installed: {e}") from e
        return attrs, {g: (d["coordinates"], d["unitcell"], float(d["philox_ctr"][0]) if "philox_ctr" in d else None)
                       for g, d in groups.items()}
    path = _npz_name(filename)  # the container written before h5min existed
    if not os.path.exists(path):
        raise FileNotFoundError(f"no restart file {filename} (or {path})")
    attrs, walkers = {}, {}
    with np.load(path, allow_pickle=False) as z:
        for k in z.files:
            if k.startswith("attr/"):
                v = z[k]
                attrs[k[5:]] = v.item() if v.shape == () else v
            else:
                g, ds = k.rsplit("/", 1)
                walkers.setdefault(g, [None, None, None])[0 if ds == "coordinates" else 1] = z[k]
    return attrs, {g: tuple(v) for g, v in walkers.items()}


def restart_attrs(args: dict, i: int, steps: dict) -> dict:
    """Root attributes of the restart file, NSio.py:74-83."""
    attrs = {}
    for j in args:
        v = args[j]
        if v is None:
            continue
        attrs[j] = v
    attrs["prev_iters"] = i + 1
    for k in ("dv_max", "dr_max", "dt_max", "dh_max", "dshear", "dstretch"):
        if k in steps and steps[k] is not None:
            attrs[k] = float(steps[k])
    return attrs


def walker_group(vrank: int, iwalker: int) -> str:
    """NSio.py:87: group of walker `iwalker` (0-based) of MPI rank `vrank`; 1-based, :04d."""
    return f"walker_{vrank}_{iwalker + 1:04d}"


def write_to_restart(args, rank_states, filename="restart.hdf5", i=0, steps=None, nwalkers=None) -> str:
    """NSio.py:69-111.  rank_states[g] = (H (nlocal,3,3), R (nlocal,N,3)[, ctr (nlocal,)]) for every GPU g, already
    gathered on the writing rank (the reference ssend/recvs lists of Atoms, NSio.py:96-110).  Groups are named per
    VIRTUAL rank -- `nwalkers` walkers each, the reference's per-MPI-rank count -- so the file is what `size` MPI ranks
    of the reference would have written (mpihans.py:171-179 reads walker_{rank}_{1..nwalkers}) and does not depend on
    how the virtual ranks were spread over GPUs."""
    walkers = {}
    g0 = 0
    for st in rank_states:
        H, R = st[0], st[1]
        ctr = st[2] if len(st) > 2 else None
        nw = int(nwalkers) if nwalkers else H.shape[0]
        for iw in range(H.shape[0]):
            g = g0 + iw
            walkers[walker_group(g // nw, g % nw)] = (R[iw], H[iw], None if ctr is None else int(ctr[iw]))
        g0 += H.shape[0]
    return save_restart(filename, restart_attrs(args, i, steps or {}), walkers)


def write_to_extxyz(H, R, filename="traj.extxyz"):
    """NSio.py:113-125: wrap the configuration into the cell and append it to the trajectory."""
    at = Atoms(f"C{len(R)}", positions=R, pbc=True, cell=H)
    at.wrap()
    write_extxyz(filename, at, append=True)


def restart_cleanup(args, traj_interval=100):
    """NSio.py:127-144: drop volumes.txt lines / trajectory frames newer than the restart point."""
    with open("volumes.txt", "r+") as f:
        lines = f.readlines()
    if (len(lines) - 1) > args["prev_iters"]:
        print("Warning, restarting from an older file, data may be overwritten/deleted")
        lines = lines[:(args["prev_iters"] + 1)]
        with open("volumes.txt", "w+") as f:
            f.writelines(lines)
        n_images = max(args["prev_iters"] // traj_interval, 1)
        if os.path.exists("traj.extxyz"):
            old = read_extxyz("traj.extxyz", f":{n_images}")
            write_extxyz("traj.extxyz", old, append=False)
