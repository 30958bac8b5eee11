import sys
sys.path.insert(0, "/root/repo")
import torch
from hans_b200 import engine, _lib
for nex in (3, 4, 5):
    pool = engine.WalkerPool(1024, 16, 6, [3,48,32,0,3,3], seed=1, nexclude=nex)
    pool.set_initial_cells(); pool.grow(); pool.mc_run(5, counters=False)
    torch.cuda.synchronize()
    lim = float(pool.vol.median().item())
    for label, mr in (("shear",[0,0,0,0,1,0]),("stretch",[0,0,0,0,0,1]),("vol",[1,0,0,0,0,0])):
        cfg = pool.cfg_with(mr)
        e0,e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pool.mc_run(1, volume_limit=lim, cfg=cfg, counters=False)
        pool.pair_tests.zero_(); _lib.diag_counters(True)
        e0.record(); att,acc = pool.mc_run(4, volume_limit=lim, cfg=cfg); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1); moves = 1024*4*pool.moves_per_sweep
        print("nex",nex,label,"us/move/walker",round(ms*1e3/(4*pool.moves_per_sweep),2),"acc",round(float(acc.sum()/att.sum()),3),"pairs/move",int(pool.pair_tests.item())/moves, "diag", _lib.diag_counters(), "moves", moves)
