#!/usr/bin/env python
"""bench.py -- headline benchmark of the nested-sampling MC walker hot path.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
    python bench.py --impl reference --steps K --warmup W    (CPU arm: the oracle port on host cores)

A "step" is one nested-sampling iteration (mpihans.py:210-245) over the resident walkers:
local/global MAXLOC, clone, and `walklength` sweeps of MC moves for every active walker.
Default workload "C3" = the configuration BASELINE.json's target is quoted on (configs[0] / configs[2], the README
input): hexamer chains nbeads=6 nchains=16, move_ratio 3,48,32,0,3,3, walklength=40, 1024 walkers per GPU -- 8192
walkers at --gpus 8; every resident walker walks each iteration (1024 virtual ranks of nwalkers=1).
Weak scaling by default (walkers per GPU fixed); --total-walkers T shards T walkers over the N GPUs instead
("scaling": "strong": configs[2]'s 8192 walkers over 2 / 4 / 8 GPUs).  The other BASELINE shapes (C2 single spheres,
C4 12x32, C5 6x128) are timed briefly after the headline and reported as `other_workloads` in the same line.
Metric: MC move attempts/s (whole job); NS iterations/s is reported beside it.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

_JSON_OUT = None  # the process's original stdout (set in main)
ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (nchains, nbeads, move_ratio, walkers per GPU, mean algorithmic pair tests per move)
    "C2": (64, 1, [1, 192, 0, 0, 3, 3], 1024, 132.0),
    "C3": (16, 6, [3, 48, 32, 0, 3, 3], 1024, 922.0),
    "C4": (32, 12, [3, 32, 32, 16, 3, 3], 512, 11235.0),
    "C5": (128, 6, [1, 384, 256, 384, 3, 3], 256, 6528.0),
}
FLOP_PER_PAIR = 45.0
WALKLENGTH = 40
SEED = 20261019


def algorithmic_pairs_per_move(nchains, nbeads, move_ratio):
    """No-early-exit inter-chain pair tests per move, weighted by the move ratio."""
    N = nchains * nbeads
    single = nbeads * (N - nbeads)
    box = (N * N - nchains * nbeads * nbeads) / 2.0
    mr = np.asarray(move_ratio, dtype=float)
    w = mr / mr.sum()
    return float(w[1] * single + w[2] * single + w[3] * single + (w[0] + w[4] + w[5]) * box)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=f, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1])); mx.append(float(p[2]))
                except ValueError:
                    continue
                for n, v in zip(names, p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            os.unlink(self.path)
        except OSError:
            pass
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(np.max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


# ---------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (oracle/) on the host cores.  TEST INFRASTRUCTURE used as the baseline.
# ---------------------------------------------------------------------------------------------------
def cpu_prepare(name, nwalk, initial_walk=20):
    from oracle import oracle as ho
    nchains, nbeads, mr, _, _ = WORKLOADS[name]
    cfg = ho.make_cfg(nchains, nbeads, mr)
    H, R = ho.grow_walkers(cfg, nwalk, SEED)
    ctrs = np.zeros(nwalk, dtype=np.uint64)
    ho.mc_run_many(cfg, H, R, np.arange(nwalk), initial_walk, 1e300, SEED, np.arange(nwalk), ctrs)
    return ho, cfg, H, R, ctrs


def cpu_native_rate(name, seconds=12.0, threads=None):
    """Variant (ii) of BASELINE.md: the oracle's native MC_run loop, one walker per OpenMP thread."""
    threads = threads or os.cpu_count()
    nwalk = max(threads * 2, 8)
    ho, cfg, H, R, ctrs = cpu_prepare(name, nwalk)
    mps = ho.moves_per_sweep(cfg.nbeads, cfg.nchains)
    vols = np.array([ho.volume(H[w]) for w in range(nwalk)])
    limit = float(vols.max())
    moves, t0 = 0, time.perf_counter()
    while True:
        ho.mc_run_many(cfg, H, R, np.arange(nwalk), WALKLENGTH, limit, SEED, np.arange(nwalk), ctrs, nthreads=threads)
        moves += nwalk * WALKLENGTH * mps
        dt = time.perf_counter() - t0
        if dt > seconds:
            break
    return moves / dt, threads, f"{nwalk} walkers x {WALKLENGTH} sweeps repeated for {dt:.1f} s on {threads} OpenMP threads"


def _shaped_worker(args):
    """Variant (i): a Python loop shaped like MC_run (MCNS.py:304-383): per move one proposal call
    and one apply call through ctypes, i.e. the reference's interpreter-bound structure."""
    name, wid, seconds = args
    import ctypes as C
    ho, cfg, H, R, ctrs = cpu_prepare(name, 1)
    L = ho.lib()
    prop = np.zeros(1, dtype=ho.PROPOSAL_DTYPE)
    dec = np.zeros(1, dtype=ho.DECISION_DTYPE)
    mps = ho.moves_per_sweep(cfg.nbeads, cfg.nchains)
    limit = ho.volume(H[0])
    accepted = np.zeros(6)
    attempted = np.zeros(6)
    m, t0 = 0, time.perf_counter()
    Hp, Rp, pp, dp = H[0].ctypes.data, R[0].ctypes.data, prop.ctypes.data, dec.ctypes.data
    while time.perf_counter() - t0 < seconds:
        for _ in range(mps):
            L.ho_gen_proposal(C.byref(cfg), SEED, wid, m, pp)
            L.ho_apply_proposal(C.byref(cfg), Hp, Rp, pp, C.c_double(limit), 0, dp)
            t = int(prop["type"][0])
            attempted[t] += 1
            accepted[t] += dec["accepted"][0]
            m += 1
    return m / (time.perf_counter() - t0)


def cpu_shaped_rate(name, seconds=8.0, procs=None):
    import multiprocessing as mp
    procs = procs or os.cpu_count()
    # spawn, not fork: the parent has live OpenMP / CUDA state that must not be inherited
    with mp.get_context("spawn").Pool(procs) as pool:
        rates = pool.map(_shaped_worker, [(name, w, seconds) for w in range(procs)])
    return float(np.sum(rates)), procs


def workload_string(name, nw):
    """config.workload: the SAME string for the GPU arm and the CPU arm (the driver compares them)."""
    nchains, nbeads, mr, _, _ = WORKLOADS[name]
    return f"{name}: nbeads={nbeads} nchains={nchains} walkers_per_gpu={nw} walklength={WALKLENGTH} move_ratio={mr}"


def run_reference(args):
    """CPU arm: nested-sampling iterations (MAXLOC, clone, walk of every walker of the sample) by the oracle's native
    loop on all host threads.  Each step is a BOUNDED SAMPLE of the workload: 2 x threads walkers instead of
    walkers_per_gpu x n_gpus (same shape, same walklength, same moves per walker)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload
    nchains, nbeads, mr, nw_default, _ = WORKLOADS[name]
    world = max(args.gpus, 1)
    nw = (args.total_walkers // world) if args.total_walkers else (args.walkers or nw_default)
    threads = os.cpu_count()
    nwalk = max(threads * 2, 8)
    ho, cfg, H, R, ctrs = cpu_prepare(name, nwalk)
    mps = ho.moves_per_sweep(nbeads, nchains)
    vols = np.array([ho.volume(H[w]) for w in range(nwalk)])
    ids = np.arange(nwalk)
    rng = np.random.default_rng(SEED)

    def step():
        # mpihans.py:211-245 on the sample: MAXLOC, clone a random walker over the max one, walk under the new ceiling
        imax = int(np.argmax(vols))
        limit = float(vols[imax])
        src = int(rng.integers(nwalk))
        H[imax], R[imax] = H[src], R[src]
        v, _, _ = ho.mc_run_many(cfg, H, R, ids, WALKLENGTH, limit, SEED, ids, ctrs, nthreads=threads)
        vols[:] = v

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    moves = args.steps * nwalk * WALKLENGTH * mps
    value = moves / dt
    sample = (f"each step = one NS iteration over a sample of {nwalk} walkers (of the workload's {nw * world}) x {WALKLENGTH} sweeps "
              f"= {nwalk * WALKLENGTH * mps} moves of the {name} shape, oracle native loop, {threads} OpenMP threads")
    line = {
        "impl": "reference", "metric": "mc_move_attempts_per_s", "value": value, "unit": "moves/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong" if args.total_walkers else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_string(name, nw), "total_walkers": nw * world,
                   "note": "oracle stand-in for hs_alkane via mpihans.py/MPI -- real hs_alkane unavailable in this image"},
        "cpu_baseline": {"value": value, "unit": "moves/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "moves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "ns_iterations_per_s": None, "gpu_launches": 0,
    }
    if args.cpu_shaped:
        try:
            srate, procs = cpu_shaped_rate(name, seconds=min(args.cpu_seconds, 8.0))
            line["cpu_baseline"]["reference_shaped"] = {
                "value": srate, "unit": "moves/s", "cores": procs,
                "sample": "per-move Python loop over the oracle's C entry points (MC_run-shaped, BASELINE.md variant i), "
                          "one process per core"}
        except Exception as e:
            line["cpu_baseline"]["reference_shaped"] = {"value": None, "sample": f"failed: {e}"}
    print(json.dumps(line), file=_JSON_OUT or sys.stdout, flush=True)


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
def kernel_name_for(pool, nbeads, tpw):
    """Which kernel hans_mc_run_batch_ws dispatches this shape to."""
    import ctypes as C
    from hans_b200 import _lib
    win = _lib.lib.hans_default_window(C.byref(pool.cfg)) if tpw == 0 else (tpw if tpw < 24 else 0)
    tpw_coop = tpw if tpw >= 32 else _lib.lib.hans_default_threads_per_walker(C.byref(pool.cfg))
    team = (_lib.lib.hans_default_team(C.byref(pool.cfg)) if tpw == 0 and not win else (tpw - 20 if tpw in (24, 28) else 0))
    if team:
        win = 0
    return (f"hb::walk_lean_kernel<NB={nbeads},W={win}> (one warp per walker)" if win
            else f"hb::walk_team_kernel<NB={nbeads if nbeads in (6, 12) else 0},NWARP={team}> (one CTA per walker, one warp per move)" if team
            else f"hb::walk_kernel<T={tpw_coop}> (one CTA per walker)") + " + hb::gen_kernel"


def ncu_traffic(name, nw):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the walk kernel, from the `ncu --set full` capture of
    this workload committed under profiles/ (profiles/r02_traffic.json, written by tools/ncu_summary.py from the
    capture); None when this workload / walker count has no capture."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        e = t.get(f"{name}:{nw}")
        return (float(e["bytes_per_launch"]), e.get("source")) if e else (None, None)
    except (OSError, ValueError, KeyError, TypeError):
        return None, None


def run_b200(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD

    from hans_b200 import engine
    from hans_b200.ns import NestedSampler

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier(group=group)
            torch.cuda.synchronize(dev)

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        return float(t.item())

    def build(name, nw, initial_walk, log_steps):
        nchains, nbeads, mr, _, _ = WORKLOADS[name]
        pool = engine.WalkerPool(nw, nchains, nbeads, mr, device=dev, seed=SEED, gid_base=rank * nw, threads_per_walker=args.tpw)
        pool.set_initial_cells()
        pool.grow()
        pool.mc_run(initial_walk, counters=False)  # perturb_initial_configs, MCNS.py:527-546
        ns = NestedSampler(pool, nwalkers=1, walklength=WALKLENGTH, rank=rank, world=world, group=group, seed=SEED,
                           traj_interval=0, log_cap=max(4096, log_steps + 8))
        ns.walk_counters = False
        return pool, ns

    def timed_region(pool, ns, steps, host=None, kernel_events=None):
        """K iterations on the device clock; returns ms (this rank).  host = (hH, hR, hvol, hlim): the e2e variant with
        the walker state uploaded from pinned host memory and the volumes read back EVERY step."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for s in range(steps):
            if host is not None:
                pool.H.copy_(host[0], non_blocking=True)
                pool.R.copy_(host[1], non_blocking=True)
                pool.refresh_volumes()
            if kernel_events is not None:
                k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ns.step(k0, k1)
                kernel_events.append((k0, k1))
            else:
                ns.step()
            if host is not None:
                host[2].copy_(pool.vol, non_blocking=True)
                host[3].copy_(ns.d_limit, non_blocking=True)
        e1.record()
        barrier()
        return e0.elapsed_time(e1)

    name = args.workload
    nchains, nbeads, mr, nw_default, _ = WORKLOADS[name]
    if args.total_walkers:
        if args.total_walkers % world:
            raise SystemExit("--total-walkers must be a multiple of the number of GPUs")
        nw = args.total_walkers // world
    else:
        nw = args.walkers or nw_default
    N = nchains * nbeads
    pool, ns = build(name, nw, args.initial_walk, args.steps + args.warmup)
    mps = pool.moves_per_sweep
    moves_per_step = world * nw * WALKLENGTH * mps
    apm = algorithmic_pairs_per_move(nchains, nbeads, mr)
    walk_kernel_name = kernel_name_for(pool, nbeads, args.tpw)
    # FP32 FFMA issue peak of this GPU (roofline denominator; MEASURED_PEAKS.json has no FP32 figure)
    ffma_peak = engine.ffma_peak_tflops()

    # ---- device-resident timing (value) --------------------------------------------------------
    clocks = ClockSampler(local)
    ns.run(args.warmup)
    barrier()
    if rank == 0:
        clocks.start()
    launches0 = pool.launches
    pool.pair_tests.zero_()
    kev = []
    ms = max_over_ranks(timed_region(pool, ns, args.steps, kernel_events=kev))
    clk = clocks.stop() if rank == 0 else {}
    launches = pool.launches - launches0
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    executed_pairs = int(pool.pair_tests.item()) / max(args.steps, 1)
    value = args.steps * moves_per_step / (ms * 1e-3)
    log, _ = ns.flush()
    vol_first, vol_last = float(log[0, 1]), float(log[-1, 1])

    # ---- end-to-end through the public API with host buffers (e2e) -----------------------------
    hH = torch.empty(pool.H.shape, dtype=torch.float64).pin_memory()
    hR = torch.empty(pool.R.shape, dtype=torch.float64).pin_memory()
    hvol = torch.empty(pool.vol.shape, dtype=torch.float64).pin_memory()
    hlim = torch.empty(1, dtype=torch.float64).pin_memory()
    hH.copy_(pool.H); hR.copy_(pool.R)
    torch.cuda.synchronize(dev)
    e2e_steps = max(args.steps // 2, 1)
    timed_region(pool, ns, min(args.warmup, 3), host=(hH, hR, hvol, hlim))
    ms_e2e = max_over_ranks(timed_region(pool, ns, e2e_steps, host=(hH, hR, hvol, hlim)))
    e2e_value = e2e_steps * moves_per_step / (ms_e2e * 1e-3)
    h2d_bytes = hH.numel() * 8 + hR.numel() * 8
    d2h_bytes = hvol.numel() * 8 + 8
    ns.flush()

    # ---- the documented drop-in call, timed: MCNS.MC_run_batch (numpy boxes in, numpy boxes + rates out) ----
    dropin = None
    if rank == 0 and not args.no_dropin:
        try:
            from hans_b200 import MCNS as NS
            a = dict(nwalkers=nw, nchains=nchains, nbeads=nbeads, bondlength=0.4, bondangle=109.7)
            NS.initialise_sim_cells(a, 1)
            NS.alk._S.H[...] = pool.H.cpu().numpy()
            NS.alk._S.R[...] = pool.R.cpu().numpy()
            boxes = range(1, nw + 1)
            lim = float(pool.vol.max().item())
            NS.MC_run_batch(a, WALKLENGTH, mr, boxes, volume_limit=lim, min_ang=45.0, min_ar=0.8)
            torch.cuda.synchronize(dev)
            reps = max(min(args.steps // 4, 10), 2)
            t0 = time.perf_counter()
            for _ in range(reps):
                NS.MC_run_batch(a, WALKLENGTH, mr, boxes, volume_limit=lim, min_ang=45.0, min_ar=0.8)
            torch.cuda.synchronize(dev)
            dt = time.perf_counter() - t0
            dropin = {"call": "hans_b200.MCNS.MC_run_batch (host numpy state in and out, INTEGRATION.md level 2)",
                      "value": reps * nw * WALKLENGTH * mps / dt, "unit": "moves/s", "calls": reps, "n_gpus": 1,
                      "ms_per_call": dt / reps * 1e3}
            NS.alk.alkane_destroy()
        except Exception as e:  # informational; never takes the line down
            dropin = {"value": None, "error": str(e)[:200]}

    # ---- dense regime: compress by nested sampling itself, then time the same step ---
    dense = None
    if args.dense_eta > 0:
        rr, dd = 0.5, 0.4   # fused-sphere chain volume, Intersecting_Sphere_Notes.ipynb
        vchain = 4 / 3 * np.pi * rr ** 3 + np.pi * (nbeads - 1) * (16 * rr ** 3 - (dd - 2 * rr) ** 2 * (dd + 4 * rr)) / 12
        # isotropic shrink steps through alkane_change_box, each kept only where no overlap appears,
        # with short constant-volume walks in between (the procedure of tests/common.py:compress, on
        # the device); nested sampling itself compresses by only 1/(K N) in ln V per iteration
        all_ids = torch.arange(nw, dtype=torch.int32, device=dev)
        relax = pool.cfg_with([0.0] + [float(x) for x in mr[1:]])
        it, eta = 0, 0.0
        while it < args.dense_max_iters:
            Hb, Rb = pool.H.clone(), pool.R.clone()
            pool.change_box(all_ids, ((args.dense_shrink - 1.0) * pool.H).cpu().numpy())
            bad = pool.check_overlap().bool()
            pool.H[bad] = Hb[bad]
            pool.R[bad] = Rb[bad]
            pool.refresh_volumes()
            pool.mc_run(2, counters=False, cfg=relax, count_pairs=False)
            it += 1
            if it % 10 == 0:
                pool.cfg.dr_max = max(0.5 * pool.cfg.dr_max, 0.05) if bad.float().mean().item() > 0.5 else pool.cfg.dr_max
                relax = pool.cfg_with([0.0] + [float(x) for x in mr[1:]])
            eta = nchains * vchain / float(pool.vol.median().item())
            if eta >= args.dense_eta:
                break
        for _ in range(4):  # adjust_mc_steps (MCNS.py:657-717) under the current ceiling
            ns.run(2)
            ns.adjust_mc_steps()
        ns.flush()
        pool.pair_tests.zero_()
        ns.walk_counters = True
        ns.run(3)
        att = ns.att.sum(0).double().cpu().numpy(); accd = ns.acc.sum(0).double().cpu().numpy()
        ns.walk_counters = False
        pool.pair_tests.zero_()
        dsteps = max(args.steps // 2, 1)
        ms_d = max_over_ranks(timed_region(pool, ns, dsteps))
        dense = {"value": dsteps * moves_per_step / (ms_d * 1e-3), "unit": "moves/s", "steps": dsteps,
                 "ms_per_step": ms_d / dsteps, "packing_fraction": eta, "compression_rounds": it,
                 "acceptance_by_type": [round(float(a / max(b, 1.0)), 3) for a, b in zip(accd, att)],
                 "step_sizes": {k: round(v, 5) for k, v in ns.steps().items()},
                 "executed_pair_tests_per_move": int(pool.pair_tests.item()) / (dsteps * nw * WALKLENGTH * mps)}
        ns.flush()
    ns.close()
    del pool, ns

    # ---- the other BASELINE shapes, briefly (same step, their own default walkers per GPU) --------
    others = []
    for oname in [x for x in args.others.split(",") if x and x != name and x in WORKLOADS]:
        try:
            onc, onb, omr, onw, _ = WORKLOADS[oname]
            opool, ons = build(oname, onw, min(args.initial_walk, 20), args.other_steps + 4)
            ons.run(1)
            opool.pair_tests.zero_()
            okev = []
            oms = max_over_ranks(timed_region(opool, ons, args.other_steps, kernel_events=okev))
            omoves = world * onw * WALKLENGTH * opool.moves_per_sweep
            okms = float(np.mean([a.elapsed_time(b) for a, b in okev]))
            oexec = int(opool.pair_tests.item()) / args.other_steps
            others.append({"workload": workload_string(oname, onw), "total_walkers": onw * world,
                           "value": args.other_steps * omoves / (oms * 1e-3), "unit": "moves/s", "steps": args.other_steps,
                           "ms_per_step": oms / args.other_steps, "ns_iterations_per_s": args.other_steps / (oms * 1e-3),
                           "kernel": kernel_name_for(opool, onb, args.tpw), "kernel_ms_per_launch": okms,
                           "executed_pair_tests_per_move": oexec / (onw * WALKLENGTH * opool.moves_per_sweep),
                           "roofline_frac": (FLOP_PER_PAIR * oexec / (okms * 1e-3) / 1e12) / ffma_peak})
            ons.close()
            del opool, ons
        except Exception as e:
            others.append({"workload": oname, "value": None, "error": str(e)[:200]})

    # ---- CPU baseline beside it (rank 0, N = 1 only) -------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            rate, threads, sample = cpu_native_rate(name, seconds=args.cpu_seconds)
            cpu = {"value": rate, "unit": "moves/s", "cores": threads, "kind": "port", "sample": sample}
            if args.cpu_shaped:
                srate, procs = cpu_shaped_rate(name, seconds=min(args.cpu_seconds, 8.0))
                cpu["reference_shaped"] = {"value": srate, "unit": "moves/s", "cores": procs,
                                           "sample": "per-move Python loop over the oracle's C entry points "
                                                     "(MC_run-shaped, BASELINE.md variant i), one process per core"}
        except Exception as e:  # the baseline must never take the GPU line down
            cpu = {"value": None, "unit": "moves/s", "cores": 0, "kind": "port", "sample": f"failed: {e}"}

    if rank == 0:
        pairs_per_launch = apm * nw * WALKLENGTH * mps
        achieved_alg = FLOP_PER_PAIR * pairs_per_launch / (kernel_ms * 1e-3) / 1e12
        achieved_exec = FLOP_PER_PAIR * executed_pairs / (kernel_ms * 1e-3) / 1e12  # (this rank's counter)
        # HBM side of the same launch: proposals (64 B per move, written by gen_kernel, read by the walk kernel)
        # + one read and one write of the walker state; shows how far from the HBM roofline the path is
        hbm_bytes = 2 * 64.0 * nw * WALKLENGTH * mps + 2.0 * nw * (72 + 24 * N + 16)
        hbm_peak = None
        try:
            hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs")
        except (OSError, ValueError):
            pass
        hbm_peak = hbm_peak or 6650.0  # GB/s; fallback of B200_PROFILING.md when the driver's file is absent
        traffic, traffic_src = ncu_traffic(name, nw)
        line = {
            "metric": "mc_move_attempts_per_s", "value": value, "unit": "moves/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if args.total_walkers else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_string(name, nw), "total_walkers": nw * world, "moves_per_iteration": moves_per_step,
                       "initial_walk_sweeps": args.initial_walk, "parallel_mode": "every resident walker walks each iteration",
                       "threads_per_walker": args.tpw or "default", "volume_first_last": [vol_first, vol_last],
                       "cache": "inputs larger than L2: every step streams a fresh proposal buffer of "
                                f"{64.0 * nw * WALKLENGTH * mps / 1e6:.0f} MB (written by the generator kernel, read once by the walk "
                                "kernel; L2 is 126 MB) -- no flush needed; the walker state itself lives in shared memory for the "
                                "whole walk"},
            "ns_iterations_per_s": args.steps / (ms * 1e-3),
            "e2e": {"value": e2e_value, "unit": "moves/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "steps": e2e_steps, "ms_per_step": ms_e2e / e2e_steps, "dropin_mc_run_batch": dropin},
            "gpu_launches": launches,
            "clocks": clk,
            # frac = EXECUTED pair tests (device counter) x 45 FLOP / kernel time / measured FFMA peak ("the
            # roofline fraction uses executed work"); the no-early-exit algorithmic count is kept beside it
            "roofline": {"bound": "fp32_alu", "achieved": achieved_exec, "peak": ffma_peak, "unit": "TFLOP/s",
                         "frac": achieved_exec / ffma_peak if ffma_peak else None, "traffic": traffic, "traffic_source": traffic_src,
                         "executed_pair_tests_per_launch": executed_pairs,
                         "executed_pair_tests_per_move": executed_pairs / (nw * WALKLENGTH * mps),
                         "algorithmic": {"pair_tests_per_launch": pairs_per_launch, "pair_tests_per_move": apm,
                                         "tflops_if_all_were_executed": achieved_alg,
                                         "note": "no-early-exit inter-chain count of SURVEY.md 8; the two pruning levels and the early "
                                                 "exit skip most of it, so this is NOT a hardware rate"},
                         "flop_per_pair_test": FLOP_PER_PAIR,
                         "hbm": {"algorithmic_bytes_per_launch": hbm_bytes, "achieved_gbs": hbm_bytes / (kernel_ms * 1e-3) / 1e9,
                                 "peak_gbs": hbm_peak, "frac": hbm_bytes / (kernel_ms * 1e-3) / 1e9 / hbm_peak,
                                 "note": "state lives in shared memory for the whole walk; the path is latency / "
                                         "instruction-fetch bound, not HBM bound (DESIGN.md section 3)"},
                         "kernel": walk_kernel_name, "kernel_ms_per_launch": kernel_ms,
                         "kernel_share_of_step": kernel_ms / (ms / args.steps),
                         "peak_source": "FFMA issue-rate probe (hans_ffma_peak, 256 FFMAs per loop trip, best of 3) run in this "
                                        "process; MEASURED_PEAKS.json holds no FP32 figure; nominal 148 SMs x 128 lanes x 2 x "
                                        "1.965 GHz = 74.4 TFLOP/s"},
            "cpu_baseline": cpu,
            "dense_regime": dense,
            "other_workloads": others,
        }
        print(json.dumps(line), file=_JSON_OUT or sys.stdout, flush=True)
    if world > 1:
        dist.barrier(group=group)
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C3", choices=sorted(WORKLOADS))
    ap.add_argument("--total-walkers", type=int, default=0,
                    help="strong scaling: this many walkers in total, sharded over the GPUs (default: walkers per GPU fixed)")
    ap.add_argument("--others", default="C2,C4,C5", help="other BASELINE shapes timed briefly after the headline ('' = none)")
    ap.add_argument("--other-steps", type=int, default=3)
    ap.add_argument("--no-dropin", action="store_true", help="skip the timing of the MCNS.MC_run_batch drop-in call")
    ap.add_argument("--walkers", type=int, default=0, help="walkers per GPU (default: the workload's)")
    ap.add_argument("--initial-walk", type=int, default=100, help="unconstrained sweeps after growth (reference: 1000)")
    ap.add_argument("--tpw", type=int, default=0, help="threads per walker (0 = library default)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--dense-eta", type=float, default=0.0,
                    help="also time the step after compressing (by NS + step adaptation) to this packing fraction; 0 = skip")
    ap.add_argument("--dense-max-iters", type=int, default=600)
    ap.add_argument("--dense-shrink", type=float, default=0.985)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--cpu-shaped", action="store_true", default=True,
                    help="also time the MC_run-shaped Python loop (BASELINE.md variant i); on by default")
    ap.add_argument("--no-cpu-shaped", dest="cpu_shaped", action="store_false")
    args = ap.parse_args()
    # stdout carries the ONE JSON line and nothing else: whatever native libraries print there while the bench runs
    # (NCCL's "NCCL version ..." banner at communicator set-up) is sent to stderr
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
