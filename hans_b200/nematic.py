"""Order-parameter analysis on the device: the reference's offline ``nematic.py`` behind the same function
names, for trajectories read with :func:`hans_b200.atoms.read_extxyz` (no ase needed).

    python -m hans_b200.nematic traj.extxyz -b 6 -c 16 [-m nematic|translational]

Reference: /root/reference/nematic.py (nematic_order_param :83-137, trans_order_param :30-55,
centre_of_masses :57-81, order_param_analysis :161-180)."""
from __future__ import annotations

import argparse as ap
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check, lib, make_cfg


def _order_params(configs, Nbeads, Nchains):
    if not torch.cuda.is_available():
        raise _lib.HansError("hans_b200 needs a CUDA device (no CPU fallback)")
    n = len(configs)
    H = np.stack([np.asarray(c.get_cell(), dtype=np.float64).reshape(3, 3) for c in configs])
    R = np.stack([np.asarray(c.get_positions(), dtype=np.float64).reshape(Nchains * Nbeads, 3) for c in configs])
    dH, dR = torch.as_tensor(H).cuda(), torch.as_tensor(R).cuda()
    p2 = torch.empty(n, dtype=torch.float64, device="cuda")
    tr = torch.empty(n, dtype=torch.float64, device="cuda")
    cfg = make_cfg(Nchains, Nbeads, [1, 1, 1, 1, 1, 1])
    check(lib.hans_order_params(C.byref(cfg), dH.data_ptr(), dR.data_ptr(), n, p2.data_ptr(), tr.data_ptr(),
                                torch.cuda.current_stream().cuda_stream))
    return p2.cpu().numpy(), tr.cpu().numpy()


def nematic_order_param(configs, Nbeads, Nchains):
    """P2 of every configuration (nematic.py:83-137)."""
    return _order_params(configs, Nbeads, Nchains)[0]


def trans_order_param(configs, Nbeads, Nchains):
    """Translational order parameter of every configuration (nematic.py:30-55)."""
    return _order_params(configs, Nbeads, Nchains)[1]


def config_picker(atoms, nem_array, lower_bound, upper_bound, filename="picked_order_param.extxyz"):
    """nematic.py:155-159."""
    from .atoms import write_extxyz
    picked = [atoms[i] for i in np.where((nem_array > lower_bound) & (nem_array < upper_bound))[0]]
    write_extxyz(filename, picked)
    return picked


def order_param_analysis(argv=None):
    """nematic.py:161-180 without the matplotlib figure."""
    from .atoms import read_extxyz
    parser = ap.ArgumentParser()
    parser.add_argument("atoms_file", metavar="f", type=str)
    parser.add_argument("-m", "--mode", type=str, default="nematic")
    parser.add_argument("-b", "--beads", type=int, required=True)
    parser.add_argument("-c", "--chains", type=int, required=True)
    parser.add_argument("-P", "--pick_configs", action="store_true")
    parser.add_argument("-l", "--lbound", type=float)
    parser.add_argument("-u", "--ubound", type=float)
    args = parser.parse_args(argv)
    configs = read_extxyz(args.atoms_file, index=":")
    fn = nematic_order_param if args.mode in ("nematic", "n") else trans_order_param
    arr = fn(configs, args.beads, args.chains)
    if args.pick_configs and (args.lbound is None or args.ubound is None):
        parser.error("--pick_configs requires --lbound and --ubound.")
    for i, v in enumerate(arr):
        print(i, v)
    if args.pick_configs:
        config_picker(configs, arr, args.lbound, args.ubound)
    return arr


if __name__ == "__main__":
    order_param_analysis()
