import sys, os, json
sys.path.insert(0, "/root/repo")
import torch
from hans_b200 import engine
for name,(nc,nb,nw) in {"C2":(64,1,1024),"C3":(16,6,1024)}.items():
    for label,mr in [("trans-only",[0,1,0,0,0,0]),("rot-only",[0,0,1,0,0,0]),("vol-only",[1,0,0,0,0,0]),("shear-only",[0,0,0,0,1,0])]:
        if nb==1 and label=="rot-only": continue
        full = [1,192,0,0,3,3] if nb==1 else [3,48,32,0,3,3]
        pool = engine.WalkerPool(nw, nc, nb, full, seed=1)
        pool.set_initial_cells(); pool.grow(); pool.mc_run(5, counters=False)
        torch.cuda.synchronize()
        lim = float(pool.vol.median().item())
        cfg = pool.cfg_with(mr)
        sweeps = 40 if "only" in label and label[0] in "tr" else 4
        e0,e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pool.mc_run(1, volume_limit=lim, cfg=cfg, counters=False)
        e0.record(); att,acc = pool.mc_run(sweeps, volume_limit=lim, cfg=cfg); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1); moves = nw*sweeps*pool.moves_per_sweep
        print(name,label,"ms",round(ms,2),"Mmoves/s",round(moves/ms/1e3,1),"ns/move/walker",round(ms*1e6/(sweeps*pool.moves_per_sweep),1),"acc",float(acc.sum()/att.sum()))
