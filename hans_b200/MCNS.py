"""Mirror of the reference's ``NesSa.MCNS`` on top of the B200 engine.

Same function names, argument meaning and return values as /root/reference/NesSa/MCNS.py so that
code written against ``from NesSa import MCNS as NS`` runs with
``from hans_b200 import MCNS as NS``.  `NS.alk` is the hs_alkane stand-in (hans_b200/alk.py).

`MC_run` is the drop-in: ONE fused kernel launch (hans_mc_run_batch) replaces the Python move loop
of MCNS.py:304-383 for the box.  The helpers around it keep the reference's names and signatures
but are written for this engine: whole-box array operations on the host mirror of the walker state
and single calls into the C ABI where the reference loops over chains and beads in Python.  The
reference's own move loop is NOT restated here -- tests/test_reference_live_cpu.py executes the
unmodified /root/reference/NesSa/MCNS.py against the CPU checker and records its call sequence,
which tests/test_gpu_reference_golden.py replays on `alk`.
"""
from __future__ import annotations

import math
import sys

import numpy as np
import torch

from . import _lib, alk
from .atoms import Atoms

move_types = ['box', 'translate', 'rotate', 'dihedral', 'shear', 'stretch']  # MCNS.py:25
ivol = 0; itrans = 1; irot = 2; idih = 3; ishear = 4; istr = 5                 # MCNS.py:27


def _box(ibox):
    """Host mirror (cell, beads) of box `ibox` (1-based), as writable arrays."""
    w = alk._w(ibox)
    alk._pool()
    return alk._S.H[w], alk._S.R[w]


def mk_ase_config(ibox, Nbeads, Nchains, scaling=3.75):
    """MCNS.py:30-53: box `ibox` as an Atoms object (ase is replaced by hans_b200.atoms.Atoms); one slice
    of the host mirror instead of a loop over chains."""
    H, R = _box(ibox)
    n = int(Nbeads) * int(Nchains)
    return Atoms("C%d" % n, positions=R[:n] * scaling, pbc=True, cell=H * scaling)


def min_aspect_ratio(ibox):
    """MCNS.py:97-118, evaluated by the device metrics kernel (hans_cell_metrics)."""
    return alk.box_min_aspect_ratio(ibox)


def min_angle(ibox):
    """MCNS.py:120-133 (radians): arccos of the largest |cos| between two cell vectors (hans_cell_metrics)."""
    return math.acos(min(1.0, float(alk._metrics(alk._w(ibox))[2])))


def _shape_proposal(kind, step_size):
    """Draws of box_shear_step (MCNS.py:148,174-175) / box_stretch_step (MCNS.py:222-227) as one proposal
    record: the vector indices from alk's uniform stream, the amplitudes from numpy's global normal stream."""
    prop = np.zeros(1, dtype=_lib.PROPOSAL_DTYPE)
    prop["type"] = kind
    i = min(int(alk.random_uniform_random() * 3), 2)
    if kind == _lib.SHEAR:
        prop["idx0"] = i
        prop["p"][0, :2] = np.random.normal(scale=step_size, size=2)
    else:
        j = min(int(alk.random_uniform_random() * 3), 2)
        prop["idx0"], prop["idx1"] = i, (j + 1) % 3 if j == i else j
        prop["p"][0, 0] = np.random.normal(scale=step_size)
    return prop


def box_shear_step(ibox, step_size, aspect_ratio_limit=0.8, angle_limit=60):
    """MCNS.py:135-207 on the device: returns (boltz, delta_H); the box is left sheared, the caller reverts."""
    return alk._shape_step(ibox, _shape_proposal(_lib.SHEAR, step_size), aspect_ratio_limit, angle_limit)


def box_stretch_step(ibox, step_size, aspect_ratio_limit=0.8, angle_limit=60):
    """MCNS.py:209-252 on the device: returns (boltz, delta_H)."""
    return alk._shape_step(ibox, _shape_proposal(_lib.STRETCH, step_size), aspect_ratio_limit, angle_limit)


def _run_fused(boxes, sweeps, move_ratio, volume_limit, dshear, dstretch, min_ang, min_ar, pressure):
    """hans_mc_run_batch on the given 0-based boxes of the alk pool: host boxes -> device, one
    launch, device -> host.  Returns (volumes, rates[n,6])."""
    import math
    pool = alk._pool()
    alk._sync_cfg()
    c = pool.cfg_with(np.asarray(move_ratio, dtype=np.float64))
    c.dshear, c.dstretch = float(dshear), float(dstretch)
    c.min_ar = float(min_ar)
    c.cos_min_ang = math.cos(float(min_ang) * math.pi / 180.0)
    c.pressure = float(pressure)
    idx = torch.as_tensor(np.asarray(boxes, dtype=np.int64))
    pool.H[idx.to(pool.device)] = torch.from_numpy(alk._S.H[boxes]).to(pool.device)
    pool.R[idx.to(pool.device)] = torch.from_numpy(alk._S.R[boxes]).to(pool.device)
    att, acc = pool.mc_run(int(sweeps), ids=boxes, volume_limit=float(volume_limit), cfg=c)
    alk._S.H[boxes] = pool.H[idx.to(pool.device)].cpu().numpy()
    alk._S.R[boxes] = pool.R[idx.to(pool.device)].cpu().numpy()
    att = att.cpu().numpy().astype(np.float64)
    acc = acc.cpu().numpy().astype(np.float64)
    rates = acc / np.where(att == 0, 1, att)  # MCNS.py:384-385
    vols = pool.vol.cpu().numpy()[boxes]
    return vols, rates


def MC_run(ns_data, sweeps, move_ratio, ibox, volume_limit=sys.float_info.max, return_ase=False,
           dshear=1.0, dstretch=1.0, min_ang=60, min_ar=0.8, pressure=0, two_phase=False):
    """MC_run (MCNS.py:276-392) as ONE fused GPU launch.  Same arguments and return value:
    (volume, acceptance_rate[6]) or the Atoms object if return_ase."""
    if two_phase:
        raise NotImplementedError("two_phase volume moves (MCNS.py:254-273) are dead code in the reference")
    vols, rates = _run_fused([int(ibox) - 1], sweeps, move_ratio, volume_limit, dshear, dstretch, min_ang, min_ar,
                             pressure)
    if return_ase:
        return mk_ase_config(ibox, ns_data["nbeads"], ns_data["nchains"], scaling=1)
    return float(vols[0]), rates[0]


def MC_run_batch(ns_data, sweeps, move_ratio, iboxes, volume_limit=sys.float_info.max, dshear=1.0, dstretch=1.0,
                 min_ang=60, min_ar=0.8, pressure=0):
    """MC_run for several boxes in one launch (the GPU analogue of one walk per MPI rank).
    Returns (volumes[n], rates[n,6])."""
    return _run_fused([int(i) - 1 for i in iboxes], sweeps, move_ratio, volume_limit, dshear, dstretch, min_ang,
                      min_ar, pressure)


def clone_walker(ibox_source, ibox_clone):
    """MCNS.py:510-524: cell and beads of one box over another (two array copies on the host mirror)."""
    (Hs, Rs), (Hd, Rd) = _box(ibox_source), _box(ibox_clone)
    Hd[...] = Hs
    Rd[...] = Rs


def perturb_initial_configs(ns_data, move_ratio, walk_length=20):
    """MCNS.py:527-546: `walk_length` unconstrained sweeps on every box (one launch for all)."""
    nwalkers = ns_data["nwalkers"]
    vols, _ = MC_run_batch(ns_data, walk_length, move_ratio, range(1, nwalkers + 1),
                           min_ang=ns_data["min_angle"], min_ar=ns_data["min_aspect_ratio"])
    return {ibox: float(vols[ibox - 1]) for ibox in range(1, nwalkers + 1)}


def import_ase_to_ibox(atoms, ibox, ns_data, scaling=1.0):
    """MCNS.py:550-580: write an Atoms object (cubic `cell` of 3 numbers or a full 3x3) into box `ibox`."""
    H, R = _box(ibox)
    cell = np.asarray(atoms.cell, dtype=np.float64)
    H[...] = (np.diag(cell) if cell.size == 3 else cell.reshape(3, 3)) * scaling
    pos = np.asarray(atoms.get_positions(), dtype=np.float64)
    R[...] = pos[:R.shape[0]] * scaling


# hs_alkane's set-up protocol (MCNS.py:595-606): (entry point, key of `args` or literal value), in call order
_SETUP = (("box_set_num_boxes", "nwalkers"), ("box_initialise", None), ("box_set_pbc", 1),
          ("alkane_set_nchains", "nchains"), ("alkane_set_nbeads", "nbeads"), ("alkane_initialise", None),
          ("box_set_isotropic", 1), ("box_set_bypass_link_cells", 1), ("box_set_use_verlet_list", 0),
          ("alkane_set_bondlength", "bondlength"), ("alkane_set_bondangle", "bondangle"))


def initialise_sim_cells(args, quiet):
    """MCNS.py:582-606: configure the engine and allocate `nwalkers` boxes on the device."""
    alk.box_set_quiet(quiet)
    for name, what in _SETUP:
        fn = getattr(alk, name)
        if what is None:
            fn()
        else:
            v = args[what] if isinstance(what, str) else what
            fn(float(v) if name.startswith("alkane_set_bond") else int(v))


def default_move_ratio(ns_data):
    """MCNS.py:610-626: the six move weights (volume, translate, rotate, dihedral, shear, stretch), from the input's
    `move_ratio` (string "a,b,c,d,e,f" or array) or the per-degree-of-freedom default.  A missing `from_restart`
    counts as 0 -- the README input lacks it."""
    given = ns_data.get("move_ratio") if hasattr(ns_data, "get") else None
    if given is not None:
        mr = np.array([float(x) for x in given.split(",")] if isinstance(given, str) else given, dtype=np.float64).ravel()
        if mr.size != 6:
            raise Exception("Move ratio array should be of length 6.")
        return mr
    nc, nb = float(ns_data["nchains"]), int(ns_data["nbeads"])
    return np.array([1.0, 3.0 * nc, 2.0 * nc * (nb >= 2), nc * max(nb - 3.0, 0.0), 3.0, 3.0])


def create_initial_configs(args, max_vol_per_atom=15):
    """MCNS.py:628-632: cubic cells of side 0.999 cbrt(max_vol_per_atom N) for every box, then growth."""
    alk._pool()
    side = 0.999 * np.cbrt(args["nbeads"] * args["nchains"] * max_vol_per_atom)
    alk._S.H[: args["nwalkers"]] = side * np.eye(3)
    populate_boxes(args)


def populate_boxes(args):
    """MCNS.py:634-644: all chains of all boxes grown by ONE kernel launch (hans_grow) instead of the reference's
    retry loop over alkane_grow_chain calls."""
    pool = alk._pool()
    alk._upload_all()
    pool.grow()
    alk._download_all()
    alk.alkane_set_nchains(args["nchains"])


class SingleComm:
    """Stand-in for an mpi4py communicator of size 1 (adjust_mc_steps takes `comm`)."""

    def Get_size(self): return 1
    def Get_rank(self): return 0

    def Allreduce(self, src, dst, op=None):
        dst[...] = src


class TorchComm:
    """The two mpi4py calls adjust_mc_steps makes (MCNS.py:673,685) on torch.distributed."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group

    def Get_size(self): return self.dist.get_world_size(self.group)
    def Get_rank(self): return self.dist.get_rank(self.group)

    def Allreduce(self, src, dst, op=None):
        backend = self.dist.get_backend(self.group)
        t = torch.as_tensor(np.asarray(src, dtype=np.float64))
        if backend == "nccl":
            t = t.cuda()
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        dst[...] = t.cpu().numpy()


def adjust_mc_steps(args, comm, move_ratio, vol_max, walklength=10, lower_bound=0.2, upper_bound=0.5,
                    min_dstep=1e-5 * np.ones(6), dv_max=10.0, dr_max=10.0, dshear=1.0, dstretch=1.0):
    """MCNS.py:657-717.  The six single-type trial walks run on copies of the chosen box in ONE
    batch (the reference walks and restores the box six times)."""
    from .ns import update_steps
    size = comm.Get_size()
    rate = np.zeros(6)
    avg_rate = np.zeros_like(rate)
    mc_box = np.random.randint(args["nwalkers"])
    backup = (alk._S.H[mc_box].copy(), alk._S.R[mc_box].copy())
    for i in range(6):
        if move_ratio[i] != 0:
            rate += MC_run(args, walklength, np.eye(6)[i], mc_box + 1, vol_max, dshear=dshear, dstretch=dstretch,
                           min_ang=args["min_angle"], min_ar=args["min_aspect_ratio"])[1]
            alk._S.H[mc_box][...] = backup[0]
            alk._S.R[mc_box][...] = backup[1]
    comm.Allreduce(rate, avg_rate)
    avg_rate = avg_rate / size
    steps = dict(dv_max=alk.alkane_get_dv_max(), dr_max=alk.alkane_get_dr_max(), dt_max=alk.alkane_get_dt_max(),
                 dh_max=alk.alkane_get_dh_max(), dshear=dshear, dstretch=dstretch)
    new = update_steps(steps, avg_rate, move_ratio, lower_bound, upper_bound, min_dstep, dv_max, dr_max)
    alk.alkane_set_dv_max(new["dv_max"])
    alk.alkane_set_dr_max(new["dr_max"])
    alk.alkane_set_dt_max(new["dt_max"])
    alk.alkane_set_dh_max(new["dh_max"])
    return avg_rate, new["dshear"], new["dstretch"]
