"""Level-2 golden fixture at the C1 shape (the README input: 16 hexamers, move_ratio 3,48,32,0,3,3, walklength 40,
K = 20 walkers on one rank).  Run in the BUILD container only (8 oracle runs in parallel, ~3 min):

    python tests/golden/make_ns_golden_c1.py

  ns_curves_c1.json   per seed: ln V at checkpoints; ensemble mean / standard error; and, per seed, the <V>(T) and
                      Cvp(T) columns of ns_analyse (oracle/ns_analyse_port.py, pinned to the reference's own ns_analyse by
                      ns_analyse_golden.json) over a T grid.  The reference's driver analyses with T_min = 0.01,
                      dT = 0.001, n_T = 1000 (mpihans.py:311-329): with E = V that grid (T <= 1.01) lies below every
                      volume a short run reaches (<V> ~ 16 T), so the grid here covers the sampled range instead.
"""
import json
import multiprocessing as mp
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

CFG = dict(nchains=16, nbeads=6, move_ratio=[3, 48, 32, 0, 3, 3], K=20, walklength=40, iterations=1600)
NSEEDS = 8
NCHECK = 20
T_GRID = [8.0, 10.0, 12.0, 15.0, 20.0, 25.0, 30.0, 40.0, 50.0]  # (at T = 60 the dominant iteration is the first one)
DOF = 16 + 3 * 16  # mpihans.py:189-196


def one(seed):
    from oracle import ns_oracle
    r = ns_oracle.run_ns(CFG["nchains"], CFG["nbeads"], CFG["move_ratio"], CFG["K"], CFG["walklength"], CFG["iterations"],
                         seed=5000 + seed)
    return r["volumes"]


def main():
    from oracle import ns_analyse_port as nap
    with mp.get_context("spawn").Pool(min(NSEEDS, os.cpu_count())) as pool:
        runs = np.array(pool.map(one, range(NSEEDS)))
    lnv = np.log(runs)
    idx = np.linspace(0, CFG["iterations"] - 1, NCHECK).astype(int)
    V = np.array([[nap.analyse(r, CFG["K"], DOF, T)["V"] for T in T_GRID] for r in runs])
    Cvp = np.array([[nap.analyse(r, CFG["K"], DOF, T)["Cvp"] for T in T_GRID] for r in runs])
    mode = np.array([[nap.analyse(r, CFG["K"], DOF, T)["mode_i"] for T in T_GRID] for r in runs])
    # compressibility vs pressure: with E = V the sums are isobaric ones with beta P = 1 / T, so Var(V) = (Cvp - dof/2) T^2 and
    # the reduced isothermal compressibility is kappa_T P = beta P Var(V) / <V> = (Cvp - dof/2) T / <V> at pressure P* = 1 / T
    kap = (Cvp - DOF / 2.0) * np.array(T_GRID)[None, :] / V
    out = dict(config=CFG, nseeds=NSEEDS, dof=DOF, checkpoints=idx.tolist(), mean_lnV=lnv[:, idx].mean(0).tolist(),
               se_lnV=(lnv[:, idx].std(0, ddof=1) / np.sqrt(NSEEDS)).tolist(), T_grid=T_GRID,
               V_mean=V.mean(0).tolist(), V_sd=V.std(0, ddof=1).tolist(), Cvp_mean=Cvp.mean(0).tolist(),
               Cvp_sd=Cvp.std(0, ddof=1).tolist(), mode_i_mean=mode.mean(0).tolist(),
               pressure=[1.0 / T for T in T_GRID], kappaP_mean=kap.mean(0).tolist(), kappaP_sd=kap.std(0, ddof=1).tolist(),
               final_volume_mean=float(runs[:, -1].mean()))
    json.dump(out, open(os.path.join(HERE, "ns_curves_c1.json"), "w"), indent=1)
    print("V(T)", np.round(V.mean(0), 1), "+-", np.round(V.std(0, ddof=1), 1))
    print("Cvp(T)", np.round(Cvp.mean(0), 1), "+-", np.round(Cvp.std(0, ddof=1), 1))
    print("kappa_T P", np.round(kap.mean(0), 3), "+-", np.round(kap.std(0, ddof=1), 3))
    print("mode iteration", mode.mean(0), "first / last volume", runs[:, 0].mean(), runs[:, -1].mean())


if __name__ == "__main__":
    main()
