#!/usr/bin/env python
"""One launch of a walk kernel variant on dense C4 / C5 walkers, for `ncu -k regex:<kernel> -s 2 -c 1`.
    python tools/cells_profile.py C4 0.45 4160 [sweeps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WORKLOADS, SEED  # noqa: E402
from hans_b200 import engine  # noqa: E402
from hans_b200.ns import NestedSampler  # noqa: E402

name, rho_target, tpw = sys.argv[1], float(sys.argv[2]), int(sys.argv[3])
sweeps = int(sys.argv[4]) if len(sys.argv) > 4 else 10
# preparation runs on another kernel family than the profiled one, so that `-k regex:` picks the profiled launches only
prep_tpw = int(sys.argv[5]) if len(sys.argv) > 5 else (24 if name == "C4" else 28)
nchains, nbeads, mr, nw, _ = WORKLOADS[name]
N = nchains * nbeads
dev = torch.device("cuda:0")
pool = engine.WalkerPool(nw, nchains, nbeads, mr, device=dev, seed=SEED, threads_per_walker=prep_tpw)
pool.set_initial_cells()
pool.grow()
pool.mc_run(20, counters=False)
ns = NestedSampler(pool, nwalkers=1, walklength=10, seed=SEED, traj_interval=0)
all_ids = torch.arange(nw, dtype=torch.int32, device=dev)
relax = pool.cfg_with([0.0] + [float(x) for x in mr[1:]])
it = 0
while float((N / pool.vol).median().item()) < rho_target and it < 600:
    Hb, Rb = pool.H.clone(), pool.R.clone()
    pool.change_box(all_ids, ((0.985 - 1.0) * pool.H).cpu().numpy())
    bad = pool.check_overlap().bool()
    pool.H[bad] = Hb[bad]
    pool.R[bad] = Rb[bad]
    pool.refresh_volumes()
    pool.mc_run(2, counters=False, cfg=relax, count_pairs=False)
    it += 1
    if it % 10 == 0 and bad.float().mean().item() > 0.5:
        pool.cfg.dr_max = max(0.5 * pool.cfg.dr_max, 0.05)
        relax = pool.cfg_with([0.0] + [float(x) for x in mr[1:]])
for _ in range(3):
    ns.run(1)
    ns.adjust_mc_steps()
ns.flush()
torch.cuda.synchronize()
pool.tpw = tpw
lim = float(pool.vol.max().item())
for rep in range(3):
    pool.pair_tests.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    att, acc = pool.mc_run(sweeps, volume_limit=lim)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    moves = nw * sweeps * pool.moves_per_sweep
    print(f"{name} rho {float((N / pool.vol).median().item()):.3f} tpw {tpw}: {moves / ms / 1e3:.1f} M moves/s, "
          f"{int(pool.pair_tests.item()) / moves:.1f} pair tests per move, acceptance "
          f"{(acc.sum(0).double() / att.sum(0).double().clamp(min=1)).cpu().numpy().round(3)}", flush=True)
