"""Nested-sampling driver: the B200 counterpart of /root/reference/mpihans.py.

    python -m hans_b200.mpihans < input.txt > log.out                      (one GPU)
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 \
        -m hans_b200.mpihans -f input.txt                                  (one process per GPU)

Same stdin `key = value` input (NSio.read_hans_file), same outputs in the run directory:
volumes.txt (mpihans.py:199-202,221), traj.extxyz every 100 iterations (mpihans.py:38,232-234),
restart.hdf5 every 50 000 iterations and at the end with one backup (mpihans.py:268-277,294-298),
ns_analyse.out when `analyse = 1` (mpihans.py:308-329).

`nwalkers` keeps the reference's meaning: walkers PER RANK.  One MPI rank of the reference becomes
one *virtual rank*; a GPU hosts `ranks_per_gpu` of them (new optional key, default 1), so the total
is K = nwalkers * n_gpus * ranks_per_gpu and every iteration walks one walker per virtual rank
(mpihans.py:232-241).  MPI collectives are replaced per SURVEY.md 8e (see hans_b200/ns.py).

Deviations from the reference, all deliberate: `from_restart` and `time` are
optional; the wall-clock stop actually fires (mpihans.py:260 compares instead of assigning); the
`dshear` input sets dshear (mpihans.py:141-144 sets dstretch); only rank 0 writes ns_analyse.out;
proposals come from Philox (`seed` key; random if absent, printed and stored in the restart).
"""
from __future__ import annotations

import argparse
import os
import sys
import time as _time

import numpy as np

from . import NSio

TRAJ_INTERVAL = 100        # mpihans.py:38
RESTART_INTERVAL = 50000   # mpihans.py:268
MC_ADJUST_WL = 10          # mpihans.py:146


def compute_dof(SimParams, move_ratio) -> int:
    """mpihans.py:189-196."""
    dof = 0
    if move_ratio[0] != 0:
        dof += SimParams["nchains"]
    if move_ratio[1] != 0:
        dof += 2 * SimParams["nchains"] if SimParams["nbeads"] <= 2 else 3 * SimParams["nchains"]
    return dof


def volumes_header(K, dof, nchains) -> str:
    """mpihans.py:202 (note the trailing space)."""
    return f'{K} {1} {dof} {False} {nchains} \n'


def volumes_line(i, vol_max) -> str:
    """mpihans.py:221."""
    return f"{i} {vol_max:.13f} {vol_max:.13f} \n"


def gen_directory(SimParams, size) -> str:
    """mpihans.py:52-58."""
    prefix = f"NeSa_{SimParams['nchains']}_{SimParams['nbeads']}mer.{SimParams['nwalkers'] * size}.{SimParams['walklength'] * size}"
    i_n = 1
    while os.path.exists(f"{prefix}.{i_n}/"):
        i_n += 1
    return f"{prefix}.{i_n}/"


def chunk_end(i, last, mc_adjust_interval, traj_interval, restart_interval, max_chunk):
    """Last iteration (inclusive) of the chunk that starts at iteration i: the host has to step in
    after an iteration that adjusts step sizes (mpihans.py:248), stages a trajectory frame (:233),
    triggers a restart (:268) or ends the run."""
    j = i
    while True:
        if j >= last or j % mc_adjust_interval == 0 or (traj_interval and j % traj_interval == 0) \
                or (j + 1) % restart_interval == 0 or (j - i + 1) >= max_chunk:
            return j
        j += 1


class _Dist:
    """torch.distributed plumbing (replaces mpi4py's COMM_WORLD, mpihans.py:17-20)."""

    def __init__(self):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.group = None
        self.dist = None
        if self.world > 1:
            import torch
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            torch.cuda.set_device(self.local)
            dist.init_process_group("nccl", device_id=torch.device(f"cuda:{self.local}"))
            self.dist, self.group = dist, dist.group.WORLD

    def bcast(self, obj, root=0):
        if self.world == 1:
            return obj
        box = [obj]
        self.dist.broadcast_object_list(box, src=root, group=self.group)
        return box[0]

    def barrier(self):
        if self.world > 1:
            self.dist.barrier(group=self.group)

    def gather_states(self, pool):
        """Every rank's (H, R, Philox counters) on every rank (restart writer, NSio.py:96-110)."""
        import torch
        if self.world == 1:
            H, R = pool.get_state()
            return [(H, R, pool.ctr.cpu().numpy())]
        gH = torch.empty((self.world,) + tuple(pool.H.shape), dtype=torch.float64, device=pool.device)
        gR = torch.empty((self.world,) + tuple(pool.R.shape), dtype=torch.float64, device=pool.device)
        gC = torch.empty((self.world,) + tuple(pool.ctr.shape), dtype=torch.int64, device=pool.device)
        self.dist.all_gather_into_tensor(gH, pool.H.contiguous(), group=self.group)
        self.dist.all_gather_into_tensor(gR, pool.R.contiguous(), group=self.group)
        self.dist.all_gather_into_tensor(gC, pool.ctr.contiguous(), group=self.group)
        gH, gR, gC = gH.cpu().numpy(), gR.cpu().numpy(), gC.cpu().numpy()
        return [(gH[r], gR[r], gC[r]) for r in range(self.world)]

    def close(self):
        if self.world > 1:
            self.dist.barrier(group=self.group)
            self.dist.destroy_process_group()


def _rotate_restart(name):
    """mpihans.py:270-273,294-297 (applied to whichever container exists)."""
    for path, bak in zip(NSio.restart_paths(name), NSio.restart_paths("restart_backup.hdf5")):
        if os.path.exists(bak):
            os.remove(bak)
        if os.path.exists(path):
            os.rename(path, bak)


def main(SimParams, comm: "_Dist | None" = None):
    import torch
    from . import MCNS as NS  # noqa: F401  (default_move_ratio)
    from .engine import WalkerPool
    from .ns import NestedSampler
    from .atoms import read_extxyz

    comm = comm or _Dist()
    rank, world = comm.rank, comm.world
    t0 = _time.time()
    torch.cuda.set_device(comm.local)
    dev = torch.device(f"cuda:{comm.local}")

    SimParams = comm.bcast(SimParams)
    rpg = int(SimParams.get("ranks_per_gpu", 1))
    size = world * rpg  # number of (virtual) ranks == the reference's comm size

    # ---- run directory (mpihans.py:41-72) ---------------------------------------------------------
    directory, from_restart = None, None
    if rank == 0:
        if "from_restart" in SimParams and int(SimParams["from_restart"]):
            print(f"Attempting to restart from {SimParams['directory']}")
            from_restart, directory = True, SimParams["directory"]
        else:
            from_restart = False
            if "directory" in SimParams:
                if SimParams["directory"] == "gen_prefix":
                    print("Generating folder(s) with appropriate prefix to for NS runs")
                    directory = gen_directory(SimParams, size)
                else:
                    directory = SimParams["directory"]
                    print(f"Using directory {directory} specified in input file.")
                os.mkdir(f"{directory}")
            else:
                print("No directory specified, using default")
                directory = "./"
    directory = comm.bcast(directory)
    from_restart = comm.bcast(from_restart)
    os.chdir(f"{directory}")

    SimParams.setdefault("prev_iters", 0)
    restart_walkers = None
    if from_restart:
        attrs, restart_walkers = NSio.open_restart(SimParams["restart_file"])  # mpihans.py:77-89
        # `ranks_per_gpu` is this implementation's own key: if the resume input names it, it wins, so that a run can be
        # resumed on another number of GPUs (the file is laid out per virtual rank); everything else comes from the file
        keep = ("restart_file", "time") + (("ranks_per_gpu",) if "ranks_per_gpu" in SimParams else ())
        for k, v in attrs.items():
            if k not in keep:
                if isinstance(v, np.floating):
                    v = float(v)
                elif isinstance(v, np.integer):
                    v = int(v)
                elif isinstance(v, np.ndarray) and v.dtype.kind in "US":
                    v = str(v)
                SimParams[k] = v
        rpg = int(SimParams.get("ranks_per_gpu", 1))
        size = world * rpg
        if len(restart_walkers) != int(SimParams["nwalkers"]) * size:
            raise ValueError(f"restart file holds {len(restart_walkers)} walkers = nwalkers {SimParams['nwalkers']} x "
                             f"{len(restart_walkers) // max(int(SimParams['nwalkers']), 1)} ranks, but this run has {world} GPU(s) x "
                             f"ranks_per_gpu {rpg}; set ranks_per_gpu so that GPUs x ranks_per_gpu matches (mpihans.py:171-179)")
        if rank == 0:
            NSio.restart_cleanup(SimParams, int(SimParams.get("traj_interval", TRAJ_INTERVAL)) or TRAJ_INTERVAL)
        comm.barrier()
    elif rank == 0:
        print("New Run")
        print("Files output to:", directory)

    if "seed" not in SimParams:
        SimParams["seed"] = comm.bcast(int.from_bytes(os.urandom(6), "little"))
    seed = int(SimParams["seed"])

    if rank == 0:
        for arg in SimParams:
            print(f"{arg:<16} {SimParams[arg]}")

    move_ratio = NS.default_move_ratio(SimParams)  # mpihans.py:106-107
    SimParams["move_ratio"] = move_ratio

    # ---- step sizes (mpihans.py:118-144) -----------------------------------------------------------
    steps = dict(dv_max=float(SimParams.get("dv_max", 2.0)), dr_max=float(SimParams.get("dr_max", 2.0)),
                 dt_max=float(SimParams.get("dt_max", 0.43)), dh_max=float(SimParams.get("dh_max", 0.4)),
                 dshear=float(SimParams.get("dshear", 1.0)), dstretch=float(SimParams.get("dstretch", 1.0)))

    nw = int(SimParams["nwalkers"])
    nlocal = nw * rpg
    pool = WalkerPool(nlocal, SimParams["nchains"], SimParams["nbeads"], move_ratio, device=dev, seed=seed,
                      gid_base=rank * nlocal, threads_per_walker=int(SimParams.get("threads_per_walker", 0)),
                      min_ar=SimParams["min_aspect_ratio"], min_ang=SimParams["min_angle"],
                      bondlength=SimParams["bondlength"], bondangle=SimParams["bondangle"], **steps)

    # ---- initial configurations (mpihans.py:148-181) -----------------------------------------------
    if not from_restart:
        if "initial_config" in SimParams:
            if rank == 0:
                print("Loading initial config")
            try:
                at = read_extxyz(f'../{SimParams["initial_config"]}')
            except OSError:
                print(f"Cannot locate {SimParams['initial_config']} in {os.getcwd()}")
                sys.exit(1)
            assert len(at) == SimParams["nchains"] * SimParams["nbeads"], "Initial config has wrong number of atoms"
            pool.set_state(np.broadcast_to(at.cell, (nlocal, 3, 3)), np.broadcast_to(at.positions, (nlocal, pool.N, 3)))
        else:
            pool.set_initial_cells()  # create_initial_configs, MCNS.py:628-644
            pool.grow()
        pool.mc_run(int(SimParams["initial_walk"]), counters=False)  # perturb_initial_configs, MCNS.py:527-546
    else:
        H = np.zeros((nlocal, 3, 3))
        R = np.zeros((nlocal, pool.N, 3))
        ctr = np.zeros(nlocal, dtype=np.int64)
        for iw in range(nlocal):  # each (virtual) rank reads its own groups, mpihans.py:171-179
            g = rank * nlocal + iw
            name = NSio.walker_group(g // nw, g % nw)
            if name not in restart_walkers:
                raise KeyError(f"{name} is not in the restart file: it holds {len(restart_walkers)} walkers, this run expects "
                               f"{nw} walkers x {size} ranks (nwalkers x GPUs x ranks_per_gpu must match, mpihans.py:171-179)")
            entry = restart_walkers[name]
            H[iw], R[iw] = entry[1], entry[0]
            if len(entry) > 2 and entry[2] is not None:
                ctr[iw] = int(entry[2])  # continue the walker's proposal stream (the reference saves no RNG state)
        pool.set_state(H, R)
        pool.ctr.copy_(torch.as_tensor(ctr))
    torch.cuda.synchronize(dev)

    K = nw * size
    mc_adjust_interval = max(K // 2, 1)  # mpihans.py:185
    dof = compute_dof(SimParams, move_ratio)
    restart_interval = int(SimParams.get("restart_interval", RESTART_INTERVAL))
    traj_interval = int(SimParams.get("traj_interval", TRAJ_INTERVAL))
    if restart_interval < 1:
        restart_interval = 1 << 62  # 0 / negative: no periodic restart files (one is still written at the end)
    traj_interval = max(traj_interval, 0)  # 0: no trajectory

    f = None
    if rank == 0:
        f = open("volumes.txt", "a+")
        if not from_restart:
            f.write(volumes_header(K, dof, SimParams["nchains"]))
    sys.stdout.flush()

    start = int(SimParams["prev_iters"])
    last = start + int(SimParams["iterations"]) - 1
    ns = NestedSampler(pool, nwalkers=nw, walklength=int(SimParams["walklength"]), rank=rank, world=world,
                       group=comm.group, seed=seed, start_iter=start, traj_interval=traj_interval,
                       mc_adjust_wl=MC_ADJUST_WL)
    time_budget = float(SimParams.get("time", float("inf")))

    def write_restart(i):
        states = comm.gather_states(pool)
        if rank == 0:
            _rotate_restart("restart.hdf5")
            to_store = {k: v for k, v in SimParams.items()}
            NSio.write_to_restart(to_store, states, filename="restart.hdf5", i=i, steps=ns.steps(), nwalkers=nw)
            print("wrote to restart")
        sys.stdout.flush()

    # ---- nested sampling loop (mpihans.py:210-281) -------------------------------------------------
    ns_t0 = _time.time()
    i = start
    interrupted = False
    while i <= last:
        j = chunk_end(i, last, mc_adjust_interval, traj_interval, restart_interval, ns.log_cap // 2)
        ns.run(j - i + 1)
        log, frames = ns.flush()
        if rank == 0:
            f.writelines(volumes_line(int(r[0]), r[1]) for r in log)
        for it, fH, fR in frames:  # the rank that owned the max walker writes it, mpihans.py:232-234
            NSio.write_to_extxyz(fH, fR, filename="traj.extxyz")
        vol_max = float(log[-1, 1])
        if j % mc_adjust_interval == 0:  # mpihans.py:248-253
            r = ns.adjust_mc_steps()
            if rank == 0:
                print(j, vol_max, r)
        if rank == 0 and (time_budget - (_time.time() - t0)) < 300.0:  # mpihans.py:257-260
            interrupted = True
        interrupted = comm.bcast(interrupted)
        if interrupted:
            if rank == 0:
                print("Out of allocated time, writing to file and exiting")
            i = j + 1
            break
        if (j + 1) % restart_interval == 0:  # mpihans.py:268-277
            write_restart(j)
        i = j + 1
    if rank == 0:
        f.close()
        print("NS RUN TIME TAKEN =", _time.time() - ns_t0)
        print("writing restart")
    write_restart(i - 1)  # mpihans.py:294-298
    sys.stdout.flush()
    comm.barrier()
    ns.close()  # (peer-memory exchange buffers, if HANS_NS_PEER=1)

    # ---- analysis hook (mpihans.py:308-329) ---------------------------------------------------------
    if SimParams["analyse"] and rank == 0:
        run_ns_analyse()
    return ns


def run_ns_analyse(filename="volumes.txt", out="ns_analyse.out"):
    """mpihans.py:308-329.  ns_analyse is pymatnest's, vendored by the reference as
    NesSa.ns_analyse_main; it is not rebuilt here -- it is imported if the reference (or
    pymatnest) is installed, else the step is skipped with a message."""
    nsa_args = {"T_min": 0.01, "dT": 0.001, "n_T": 1000, "kB": 1.0, "accurate_sum": False, "verbose": False,
                "profile": False, "skip": 0, "line_end": None, "interval": 1, "delta_pressure": 0.0,
                "dump_terms": -1.0, "entropy": False, "quiet": False, "kolmogorov_smirnov_test": False, "files": filename}
    try:
        from NesSa import ns_analyse_main  # the reference's own module, if importable
    except ImportError:
        print("analyse = 1 but NesSa.ns_analyse_main is not importable; run ns_analyse on volumes.txt separately")
        return False
    from contextlib import redirect_stdout
    print("running ns_analyse")
    with open(out, "w") as fo, redirect_stdout(fo):
        ns_analyse_main.analyse(nsa_args)
    return True


def cli(argv=None):
    ap = argparse.ArgumentParser(description="hans nested sampling on B200 (input on stdin or -f)")
    ap.add_argument("-f", "--input_file", type=str, default=None)
    a = ap.parse_args(argv)
    comm = _Dist()
    SimParams = None
    if a.input_file is not None:
        SimParams = NSio.read_hans_file(a.input_file)
    elif comm.rank == 0:
        SimParams = NSio.read_hans_file()
    prof = None
    if SimParams and SimParams.get("profile"):  # mpihans.py:334-343
        import cProfile
        prof = cProfile.Profile()
        prof.enable()
    main(SimParams, comm)
    if prof is not None:
        prof.disable()
        prof.dump_stats(f"mpihans{comm.rank}.profile")
    comm.close()


if __name__ == "__main__":
    cli()
