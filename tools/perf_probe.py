"""Quick throughput probe of hans_mc_run_batch for several shapes / threads-per-walker (dev tool)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hans_b200 import engine

CASES = {
    "C2": (64, 1, [1, 192, 0, 0, 3, 3], 1024),
    "C3_1024": (16, 6, [3, 48, 32, 0, 3, 3], 1024),
    "C3_1300": (16, 6, [3, 48, 32, 0, 3, 3], 1300),
    "C3_1536": (16, 6, [3, 48, 32, 0, 3, 3], 1536),
    "C3_2048": (16, 6, [3, 48, 32, 0, 3, 3], 2048),
    "C2_2048": (64, 1, [1, 192, 0, 0, 3, 3], 2048),
    "C3_4096": (16, 6, [3, 48, 32, 0, 3, 3], 4096),
    "C3_8192": (16, 6, [3, 48, 32, 0, 3, 3], 8192),
    "C2_4096": (64, 1, [1, 192, 0, 0, 3, 3], 4096),
    "C2_8192": (64, 1, [1, 192, 0, 0, 3, 3], 8192),
    "C4_512": (32, 12, [3, 32, 32, 16, 3, 3], 512),
    "C5_256": (128, 6, [1, 384, 256, 384, 3, 3], 256),
}
which = sys.argv[1].split(",") if len(sys.argv) > 1 else list(CASES)
tpws = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [32, 64, 128]
walk = int(sys.argv[3]) if len(sys.argv) > 3 else 40
ws_mb = int(sys.argv[4]) if len(sys.argv) > 4 else 1024
for name in which:
    nc, nb, mr, nw = CASES[name]
    for tpw in tpws:
        pool = engine.WalkerPool(nw, nc, nb, mr, seed=1, threads_per_walker=tpw, workspace_mb=ws_mb)
        pool.set_initial_cells(); pool.grow()
        pool.mc_run(5, counters=False)
        pool._ws = torch.empty(min(nw * walk * pool.moves_per_sweep * 64, pool.workspace_bytes), dtype=torch.uint8, device=pool.device) if pool.workspace_bytes else None
        torch.cuda.synchronize()
        lim = float(pool.vol.median().item())
        pool.pair_tests.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        att, acc = pool.mc_run(walk, volume_limit=lim)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        moves = nw * walk * pool.moves_per_sweep
        pt = int(pool.pair_tests.item())
        rate = (acc.sum(0).double() / att.sum(0).clamp(min=1).double()).cpu().numpy().round(3)
        print(json.dumps(dict(case=name, tpw=tpw, ws_mb=ws_mb, ms=round(ms, 2), moves_per_s=round(moves / ms * 1e3 / 1e6, 2), unit="M/s",
                              pair_tests_per_s=round(pt / ms * 1e3 / 1e9, 2), pairs_per_move=round(pt / moves, 1), acc=list(rate))), flush=True)
