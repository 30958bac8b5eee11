#!/usr/bin/env python
"""Generates tests/golden/ref_live_replay.npz and tests/golden/ref_live_trace.json by EXECUTING the
unmodified reference /root/reference/NesSa/MCNS.py in the build container (tests/ref_live.py).

    python tests/golden/make_ref_live_golden.py          (needs /root/reference; run from the repo root)

/root/reference does not exist on the GPU box, so the vectors travel as fixtures:
  ref_live_replay.npz   per case: start state, the explicit proposals, the ceiling, and what the REFERENCE's MC_run
                        (MCNS.py:276-392) made of them -- final cell / beads (with numpy's exp and with the
                        deterministic exp, see tests/test_reference_live_cpu.py), returned volume and rates.
                        tests/test_gpu_reference_golden.py replays the proposals through hans_replay on the B200.
  ref_live_trace.json   per case: the sequence of hs_alkane entry-point calls the reference made (arguments, return
                        values, and the writes it did through the chain views), replayed call by call on hans_b200.alk.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ref_live  # noqa: E402
from common import SHAPES  # noqa: E402
from oracle import oracle as ho  # noqa: E402

# (shape, seed, sweeps, ceiling factor or None, keep a call trace)
CASES = [("tiny1", 3, 6, None, True), ("tiny", 4, 5, 1.01, True), ("p7", 7, 2, None, True), ("t3", 8, 3, 1.01, True),
         ("C1", 5, 3, 1.005, False), ("C2", 6, 4, 1.002, False), ("h15", 9, 2, 1.003, False), ("dimer", 10, 2, 1.004, False),
         ("C4s", 12, 1, 1.002, False)]
STEP_KW = dict(dshear=0.4, dstretch=0.08)


def start_state(shape, seed):
    nchains, nbeads, mr = SHAPES[shape]
    cfg = ho.make_cfg(nchains, nbeads, mr, **STEP_KW)
    H, R = ho.grow_walkers(cfg, 1, seed)
    ctrs = np.zeros(1, dtype=np.uint64)
    ho.mc_run_many(cfg, H, R, [0], 2, 1e300, seed + 1, [0], ctrs)
    return cfg, mr, H[0].copy(), R[0].copy()


def main():
    plain = ref_live.LiveReference()
    det = ref_live.LiveReference(np_overrides={"exp": (lambda x: ho.det_exp(float(x)))})
    arrays, traces = {}, {}
    for shape, seed, sweeps, lf, keep in CASES:
        cfg, mr, H0, R0 = start_state(shape, seed)
        mps = ho.moves_per_sweep(cfg.nbeads, cfg.nchains)
        props = ho.gen_proposals(cfg, 1000 + seed, 0, 0, sweeps * mps)
        limit = sys.float_info.max if lf is None else ho.volume(H0) * lf
        out = {}
        for tag, L in (("np", plain), ("det", det)):
            L.setup(cfg.nchains, cfg.nbeads, H0, R0)
            vol, rates = L.mc_run(cfg, props, move_ratio=mr, volume_limit=limit, min_ang=45.0, bump_stretch=True,
                                  trace=(keep and tag == "np"))
            assert L.drained()
            out[tag] = (L.alk.H[0].copy(), L.alk.R[0].copy(), vol, rates)
            if keep and tag == "np":
                traces[shape] = dict(nchains=cfg.nchains, nbeads=cfg.nbeads, H0=H0.tolist(), R0=R0.tolist(),
                                     steps=dict(dv_max=2.0, dr_max=2.0, dt_max=0.43, dh_max=0.4),
                                     calls=[[n, a, r] for n, a, r in L.alk.trace],
                                     H_final=L.alk.H[0].tolist(), R_final=L.alk.R[0].tolist())
        k = shape + "."
        arrays[k + "shape"] = np.array([cfg.nchains, cfg.nbeads, sweeps])
        arrays[k + "move_ratio"] = np.asarray(mr, dtype=np.float64)
        arrays[k + "H0"], arrays[k + "R0"] = H0, R0
        arrays[k + "props"] = props.view(np.uint8)
        arrays[k + "limit"] = np.array([limit])
        for tag in ("np", "det"):
            arrays[k + "H_" + tag], arrays[k + "R_" + tag] = out[tag][0], out[tag][1]
            arrays[k + "vol_" + tag] = np.array([out[tag][2]])
            arrays[k + "rates_" + tag] = np.asarray(out[tag][3])
        print(f"{shape}: {len(props)} moves, rates {np.round(out['np'][3], 3)}, bit-equal np/det "
              f"{np.array_equal(out['np'][1], out['det'][1])}")
    np.savez_compressed(os.path.join(HERE, "ref_live_replay.npz"), **arrays)
    with open(os.path.join(HERE, "ref_live_trace.json"), "w") as f:
        json.dump(dict(step_kw=STEP_KW, min_ar=0.8, min_ang=45.0, cases=traces), f)
    print("wrote", os.path.getsize(os.path.join(HERE, "ref_live_replay.npz")), "+",
          os.path.getsize(os.path.join(HERE, "ref_live_trace.json")), "bytes")


if __name__ == "__main__":
    main()
