"""Mirror of the reference's ``NesSa.MCNS`` on top of the B200 engine.

Same function names, argument meaning and return values as /root/reference/NesSa/MCNS.py so that
code written against ``from NesSa import MCNS as NS`` runs with
``from hans_b200 import MCNS as NS``.  `NS.alk` is the hs_alkane stand-in (hans_b200/alk.py).

Two ways to do a walk:
  * MC_run                -- the drop-in: ONE fused kernel launch (hans_mc_run_batch) replaces the
                             Python move loop of MCNS.py:304-383 for the box.
  * MC_run_python_loop    -- the reference's own per-move control flow (MCNS.py:304-383) driven
                             through the alk entry points, kept for semantic cross-checks.
"""
from __future__ import annotations

import ctypes as C
import sys

import numpy as np
import torch

from . import _lib, alk
from .atoms import Atoms

move_types = ['box', 'translate', 'rotate', 'dihedral', 'shear', 'stretch']  # MCNS.py:25
ivol = 0; itrans = 1; irot = 2; idih = 3; ishear = 4; istr = 5                 # MCNS.py:27


def mk_ase_config(ibox, Nbeads, Nchains, scaling=3.75):
    """MCNS.py:30-53: the box as an Atoms object (ase is replaced by hans_b200.atoms.Atoms)."""
    model_positions = np.empty([Nchains * Nbeads, 3])
    cell_vectors = alk.box_get_cell(int(ibox))
    for ichain in range(0, Nchains):
        model_positions[Nbeads * ichain:Nbeads * ichain + Nbeads] = alk.alkane_get_chain(int(ichain + 1), int(ibox))
    confstring = "C" + str(Nbeads * Nchains)
    return Atoms(confstring, positions=model_positions * scaling, pbc=True, cell=cell_vectors * scaling)


def min_aspect_ratio(ibox):
    """MCNS.py:97-118."""
    vol = alk.box_compute_volume(int(ibox))
    cell = alk.box_get_cell(int(ibox)).copy()
    best = sys.float_info.max
    for i in range(3):
        vi = cell[i, :]
        vnorm_hat = np.cross(cell[(i + 1) % 3, :], cell[(i + 2) % 3, :])
        vnorm_hat = vnorm_hat / (np.sqrt(np.dot(vnorm_hat, vnorm_hat)))
        best = min(best, abs(np.dot(vnorm_hat, vi)))
    return best / np.cbrt(vol)


def min_angle(ibox):
    """MCNS.py:120-133 (radians)."""
    cell = alk.box_get_cell(int(ibox)).copy()
    best = sys.float_info.max
    for i in range(3):
        vc1 = cell[(i + 1) % 3, :]
        vc2 = cell[(i + 2) % 3, :]
        dot_prod = np.dot(vc1, vc2) / np.sqrt((np.dot(vc1, vc1)) * (np.dot(vc2, vc2)))
        if dot_prod < 0:
            dot_prod *= -1
        best = min(best, np.arccos(dot_prod))
    return best


def box_shear_step(ibox, step_size, aspect_ratio_limit=0.8, angle_limit=60):
    """MCNS.py:135-207: returns (boltz, delta_H); the box is left sheared, the caller reverts."""
    rnd_vec_ind = int(np.floor(alk.random_uniform_random() * 3))
    other_vec_ind = list(range(3))
    other_vec_ind.remove(rnd_vec_ind)
    orig_cell_copy = alk.box_get_cell(int(ibox)).copy()
    v1 = orig_cell_copy[other_vec_ind[0], :].copy()
    v2 = orig_cell_copy[other_vec_ind[1], :].copy()
    v1 /= np.sqrt(np.dot(v1, v1))
    v2 -= v1 * np.dot(v1, v2)
    v2 /= np.sqrt(np.dot(v2, v2))
    if np.isnan(np.sum(v1)) or np.isnan(np.sum(v2)):
        raise FloatingPointError("box_shear_step: degenerate cell")  # the reference prints and sys.exit()s
    rv1 = np.random.normal(scale=step_size)
    rv2 = np.random.normal(scale=step_size)
    new_cell = orig_cell_copy.copy()
    new_cell[rnd_vec_ind, :] += rv1 * v1 + rv2 * v2
    delta_H = new_cell - orig_cell_copy
    alk.alkane_change_box(int(ibox), delta_H)
    angle_limit_rad = angle_limit * np.pi / 180
    if alk.box_min_aspect_ratio(ibox) < aspect_ratio_limit:
        boltz = 0
    elif min_angle(ibox) < angle_limit_rad:
        boltz = 0
    elif alk.alkane_check_chain_overlap(int(ibox)):
        boltz = 0
    else:
        boltz = 1
    return boltz, delta_H


def box_stretch_step(ibox, step_size, aspect_ratio_limit=0.8, angle_limit=60):
    """MCNS.py:209-252."""
    cell = alk.box_get_cell(int(ibox))
    new_cell = cell.copy()
    rnd_v1_ind = int(np.floor(alk.random_uniform_random() * 3))
    rnd_v2_ind = int(np.floor(alk.random_uniform_random() * 3))
    if rnd_v1_ind == rnd_v2_ind:
        rnd_v2_ind = (rnd_v2_ind + 1) % 3
    rv = np.random.normal(scale=step_size)
    new_cell[rnd_v1_ind] *= np.exp(rv)
    new_cell[rnd_v2_ind] *= np.exp(-rv)
    delta_H = new_cell - cell
    alk.alkane_change_box(int(ibox), delta_H)
    angle_limit_rad = angle_limit * np.pi / 180
    if alk.box_min_aspect_ratio(ibox) < aspect_ratio_limit:
        boltz = 0
    elif min_angle(ibox) < angle_limit_rad:
        boltz = 0
    elif alk.alkane_check_chain_overlap(int(ibox)):
        boltz = 0
    else:
        boltz = 1
    return boltz, delta_H


def _moves_per_sweep(nbeads, nchains):
    return _lib.moves_per_sweep(nbeads, nchains)  # MCNS.py:296-301


def MC_run_python_loop(ns_data, sweeps, move_ratio, ibox, volume_limit=sys.float_info.max, return_ase=False,
                       dshear=1.0, dstretch=1.0, min_ang=60, min_ar=0.8, pressure=0):
    """The reference's move loop, MCNS.py:276-392, statement for statement, on the alk stand-in."""
    moves_accepted = np.zeros(6)
    moves_attempted = np.zeros(6)
    nbeads = ns_data["nbeads"]
    nchains = ns_data["nchains"]
    move_prob = np.cumsum(move_ratio) / np.sum(move_ratio)
    moves_per_sweep = _moves_per_sweep(nbeads, nchains)
    for _ in range(int(sweeps)):
        for _ in range(moves_per_sweep):
            ichain = int(np.floor(alk.random_uniform_random() * nchains))
            current_chain = alk.alkane_get_chain(int(ichain + 1), int(ibox))
            backup_chain = current_chain.copy()
            xi = np.random.random()
            if xi < move_prob[ivol]:
                itype = ivol
                boltz = alk.alkane_box_resize(pressure, int(ibox), 0)
            elif xi < move_prob[itrans]:
                itype = itrans
                boltz = alk.alkane_translate_chain(int(ichain + 1), int(ibox))
            elif xi < move_prob[irot]:
                itype = irot
                boltz, quat = alk.alkane_rotate_chain(int(ichain + 1), int(ibox), 0)
            elif xi < move_prob[idih]:
                itype = idih
                boltz, bead1, angle = alk.alkane_bond_rotate(int(ichain + 1), int(ibox), 1)
            elif xi < move_prob[ishear]:
                itype = ishear
                boltz, delta_H = box_shear_step(ibox, dshear, min_ar, min_ang)
            else:
                itype = istr
                boltz, delta_H = box_stretch_step(ibox, dstretch, min_ar, min_ang)
            moves_attempted[itype] += 1
            if itype == ivol:
                new_volume = alk.box_compute_volume(int(ibox))
                if (np.random.random() < boltz) and (new_volume - volume_limit) < sys.float_info.epsilon:
                    moves_accepted[itype] += 1
                else:
                    alk.alkane_box_resize(pressure, int(ibox), 1)
            elif itype == ishear or itype == istr:
                if boltz:
                    moves_accepted[itype] += 1
                else:
                    alk.alkane_change_box(int(ibox), -1.0 * delta_H)
            else:
                if np.random.random() < boltz:
                    moves_accepted[itype] += 1
                else:
                    for ibead in range(nbeads):
                        current_chain[ibead] = backup_chain[ibead]
    moves_attempted = np.where(moves_attempted == 0, 1, moves_attempted)
    rate = moves_accepted / moves_attempted
    if return_ase:
        return mk_ase_config(ibox, nbeads, nchains, scaling=1)
    return alk.box_compute_volume(int(ibox)), rate


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
    """MCNS.py:510-524."""
    nbeads = alk.alkane_get_nbeads()
    nchains = alk.alkane_get_nchains()
    cell = alk.box_get_cell(ibox_source)
    alk.box_set_cell(ibox_clone, cell)
    for ichain in range(1, nchains + 1):
        original_chain = alk.alkane_get_chain(ichain, ibox_source)
        clone_chain = alk.alkane_get_chain(ichain, ibox_clone)
        for ibead in range(nbeads):
            clone_chain[ibead][:] = original_chain[ibead][:]


def perturb_initial_configs(ns_data, move_ratio, walk_length=20):
    """MCNS.py:527-546: `walk_length` unconstrained sweeps on every box (one launch for all)."""
    nwalkers = ns_data["nwalkers"]
    vols, _ = MC_run_batch(ns_data, walk_length, move_ratio, range(1, nwalkers + 1),
                           min_ang=ns_data["min_angle"], min_ar=ns_data["min_aspect_ratio"])
    return {ibox: float(vols[ibox - 1]) for ibox in range(1, nwalkers + 1)}


def import_ase_to_ibox(atoms, ibox, ns_data, scaling=1.0):
    """MCNS.py:550-580."""
    try:
        nbeads = ns_data.parameters.nbeads
        nchains = ns_data.parameters.nchains
    except AttributeError:
        nbeads = ns_data["nbeads"]
        nchains = ns_data["nchains"]
    cell_vectors = np.array(atoms.cell, dtype=np.float64)
    if cell_vectors.size == 3:
        cell_vectors = cell_vectors * np.eye(3)
    alk.box_set_cell(int(ibox), cell_vectors * scaling)
    positions = atoms.get_positions()
    for ichain in range(1, nchains + 1):
        chain = alk.alkane_get_chain(ichain, int(ibox))
        for ibead in range(nbeads):
            chain[ibead][:] = positions[(ichain - 1) * nbeads + ibead][:] * scaling


def initialise_sim_cells(args, quiet):
    """MCNS.py:582-606."""
    alk.box_set_quiet(quiet)
    alk.box_set_num_boxes(args["nwalkers"])
    alk.box_initialise()
    alk.box_set_pbc(1)
    alk.alkane_set_nchains(int(args["nchains"]))
    alk.alkane_set_nbeads(int(args["nbeads"]))
    alk.alkane_initialise()
    alk.box_set_isotropic(1)
    alk.box_set_bypass_link_cells(1)
    alk.box_set_use_verlet_list(0)
    alk.alkane_set_bondlength(float(args["bondlength"]))
    alk.alkane_set_bondangle(float(args["bondangle"]))


def default_move_ratio(ns_data):
    """MCNS.py:610-626 (a missing `from_restart` counts as 0 -- the README input lacks it)."""
    if "move_ratio" in ns_data:
        mr = ns_data["move_ratio"]
        if isinstance(mr, str):
            move_ratio = [float(i) for i in mr.split(',')]
        else:
            move_ratio = list(np.asarray(mr, dtype=np.float64))
        if len(move_ratio) != 6:
            raise Exception("Move ratio array should be of length 6.")
        return np.array(move_ratio, dtype=np.float64)
    move_ratio = np.zeros(6)
    move_ratio[ivol] = 1
    move_ratio[itrans] = 3.0 * ns_data["nchains"]
    move_ratio[irot] = (2.0 * ns_data["nchains"]) if ns_data["nbeads"] >= 2 else 0
    move_ratio[idih] = 1.0 * max(((ns_data["nbeads"] - 3.0) * (ns_data["nchains"]), 0))
    move_ratio[ishear] = 3
    move_ratio[istr] = 3
    return np.array(move_ratio)


def create_initial_configs(args, max_vol_per_atom=15):
    """MCNS.py:628-632."""
    cell_matrix = 0.999 * np.eye(3) * np.cbrt(args["nbeads"] * args["nchains"] * max_vol_per_atom)
    for ibox in range(1, args["nwalkers"] + 1):
        alk.box_set_cell(int(ibox), cell_matrix)
    populate_boxes(args)


def populate_boxes(args):
    """MCNS.py:634-644, all boxes grown by one kernel launch (hans_grow)."""
    pool = alk._pool()
    alk._upload_all()
    pool.grow()
    alk._download_all()
    alk.alkane_set_nchains(args["nchains"])


def populate_boxes_python_loop(args):
    """MCNS.py:634-644 statement for statement (one alkane_grow_chain call per attempt)."""
    ncopy = args["nchains"]
    for ibox in range(1, args["nwalkers"] + 1):
        for ichain in range(1, ncopy + 1):
            rb_factor = 0
            alk.alkane_set_nchains(ichain)
            while rb_factor == 0:
                rb_factor, ifail = alk.alkane_grow_chain(ichain, int(ibox), 1)
                if ifail != 0:
                    rb_factor = 0


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
