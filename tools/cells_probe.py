#!/usr/bin/env python
"""Density ladder: moves/s of the walk kernels on C4 / C5 sized walkers from the dilute start to the dense regime
(which kernel should take a walker at which bead density; the ladders behind profiles/r02_cells_probe.md)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WORKLOADS, WALKLENGTH, SEED  # noqa: E402
from hans_b200 import engine  # noqa: E402
from hans_b200.ns import NestedSampler  # noqa: E402

CELLS = 4096


def main():
    names = sys.argv[1].split(",") if len(sys.argv) > 1 else ["C4", "C5"]
    variants = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else None
    dev = torch.device("cuda:0")
    for name in names:
        nchains, nbeads, mr, nw, _ = WORKLOADS[name]
        N = nchains * nbeads
        pool = engine.WalkerPool(nw, nchains, nbeads, mr, device=dev, seed=SEED)
        pool.set_initial_cells()
        pool.grow()
        pool.mc_run(20, counters=False)
        ns = NestedSampler(pool, nwalkers=1, walklength=WALKLENGTH, seed=SEED, traj_interval=0)
        mps = pool.moves_per_sweep
        tvars = variants or ([24, 28, 128] if N < 700 else [28, 128]) + [CELLS | 64, CELLS | 128, CELLS | 256]
        all_ids = torch.arange(nw, dtype=torch.int32, device=dev)
        rungs = [float(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0.0, 0.12, 0.2, 0.3, 0.4, 0.55, 0.7]
        for rho_target in rungs:
            it = 0
            relax = pool.cfg_with([0.0] + [float(x) for x in mr[1:]])
            while float((N / pool.vol).median().item()) < rho_target and it < 500:
                Hb, Rb = pool.H.clone(), pool.R.clone()
                pool.change_box(all_ids, ((0.985 - 1.0) * pool.H).cpu().numpy())
                bad = pool.check_overlap().bool()
                pool.H[bad] = Hb[bad]
                pool.R[bad] = Rb[bad]
                pool.refresh_volumes()
                pool.tpw = 128
                pool.mc_run(2, counters=False, cfg=relax, count_pairs=False)
                it += 1
                if it % 10 == 0 and bad.float().mean().item() > 0.5:
                    pool.cfg.dr_max = max(0.5 * pool.cfg.dr_max, 0.05)
                    relax = pool.cfg_with([0.0] + [float(x) for x in mr[1:]])
            pool.tpw = 128
            for _ in range(3):
                ns.run(1)
                ns.adjust_mc_steps()
            ns.flush()
            rho = float((N / pool.vol).median().item())
            row = {"workload": name, "rho_median": round(rho, 4), "steps": {k: round(v, 4) for k, v in ns.steps().items()}}
            for v in tvars:
                pool.tpw = v
                try:
                    ns.run(1)
                    torch.cuda.synchronize()
                    pool.pair_tests.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    ns.run(2)
                    e1.record()
                    torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1) / 2
                    row[f"tpw{v & 4095}{'c' if v & CELLS else ''}"] = {
                        "Mmoves_s": round(nw * WALKLENGTH * mps / ms / 1e3, 1),
                        "pairs_per_move": round(int(pool.pair_tests.item()) / (2 * nw * WALKLENGTH * mps), 1)}
                except Exception as e:
                    row[f"tpw{v}"] = str(e)[:80]
                ns.flush()
            print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
