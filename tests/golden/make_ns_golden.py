"""Generates the level-2 (statistical) golden fixtures.  Run in the BUILD container only:

    python tests/golden/make_ns_golden.py

  ns_curves.json          mean and standard error of ln V at checkpoints over independent seeds of
                          the CPU oracle's nested sampling (oracle/ns_oracle.py), two small systems
  volumes_sysB_seed0.txt  one oracle run in the reference's volumes.txt format (mpihans.py:202,221)
  ns_analyse_golden.json  what THE REFERENCE'S OWN ns_analyse (imported from /root/reference/NesSa)
                          prints for that file at a few temperatures -- pins oracle/ns_analyse_port.py
"""
import contextlib
import io
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ns_oracle  # noqa: E402

SYSTEMS = {
    # name: nchains, nbeads, move_ratio, K, walklength, iterations, dof
    "A_spheres16": dict(nchains=16, nbeads=1, move_ratio=[1, 48, 0, 0, 3, 3], K=64, walklength=20, iterations=2400),
    "B_tetramers8": dict(nchains=8, nbeads=4, move_ratio=[1, 24, 16, 8, 3, 3], K=32, walklength=10, iterations=800),
}
NSEEDS = 16
NCHECK = 24


def main():
    curves = {}
    for name, s in SYSTEMS.items():
        runs = []
        for seed in range(NSEEDS):
            r = ns_oracle.run_ns(s["nchains"], s["nbeads"], s["move_ratio"], s["K"], s["walklength"], s["iterations"],
                                 seed=1000 + seed)
            runs.append(np.log(r["volumes"]))
            print(name, seed, r["volumes"][0], r["volumes"][-1], r["steps"], flush=True)
        runs = np.array(runs)
        idx = np.linspace(0, s["iterations"] - 1, NCHECK).astype(int)
        curves[name] = dict(config=s, nseeds=NSEEDS, checkpoints=idx.tolist(), mean_lnV=runs[:, idx].mean(0).tolist(),
                            se_lnV=(runs[:, idx].std(0, ddof=1) / np.sqrt(NSEEDS)).tolist(),
                            sd_lnV=runs[:, idx].std(0, ddof=1).tolist())
        if name == "B_tetramers8":
            vols = np.exp(runs[0])
            dof = 8 + 3 * 8  # mpihans.py:189-196
            with open(os.path.join(HERE, "volumes_sysB_seed0.txt"), "w") as f:
                f.write(f'{s["K"]} {1} {dof} {False} {s["nchains"]} \n')
                for i, v in enumerate(vols):
                    f.write(f"{i} {v:.13f} {v:.13f} \n")
    json.dump(curves, open(os.path.join(HERE, "ns_curves.json"), "w"), indent=1)

    # the reference's own ns_analyse on that file
    sys.path.insert(0, "/root/reference")
    from NesSa import ns_analyse_main
    Ts = [1.0, 5.0, 10.0, 15.0, 20.0, 30.0, 50.0]
    rows = []
    for T in Ts:
        args = {"T_min": T, "dT": 0.001, "n_T": 1, "kB": 1.0, "accurate_sum": False, "verbose": False, "profile": False,
                "skip": 0, "line_end": None, "interval": 1, "delta_pressure": 0.0, "dump_terms": -1.0, "entropy": False,
                "quiet": True, "kolmogorov_smirnov_test": False, "files": os.path.join(HERE, "volumes_sysB_seed0.txt")}
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            ns_analyse_main.analyse(args)
        line = [l for l in buf.getvalue().splitlines() if not l.startswith("#")][0].split()
        # columns: T log_Z F U Cvp low mode high low_it mode_it high_it Z_fract V thermal_exp
        rows.append(dict(T=float(line[0]), log_Z=float(line[1]), U=float(line[3]), Cvp=float(line[4]), mode_i=int(line[6]),
                         V=float(line[12])))
    json.dump(dict(file="volumes_sysB_seed0.txt", rows=rows), open(os.path.join(HERE, "ns_analyse_golden.json"), "w"), indent=1)
    print(rows)


if __name__ == "__main__":
    main()
