"""Serial nested sampling driven by the CPU oracle (TEST INFRASTRUCTURE; level-2 parity).

A plain restatement of the NS loop of /root/reference/mpihans.py:210-253 for ONE rank
(size = 1: only the cloned walker walks), with the step-size rule of MCNS.py:657-717.
Used to produce tests/golden/ns_curves.json and as the statistical reference for the GPU runs.
"""
from __future__ import annotations

import numpy as np

from . import oracle as ho

STEP_NAMES = ["dv_max", "dr_max", "dt_max", "dh_max", "dshear", "dstretch"]
STEP_MIN = np.array([1e-10, 1e-10, 1e-10, 1e-10, 1e-5, 1e-5])  # mpihans.py:27-33


def update_steps(steps, avg_rate, move_ratio, lower=0.2, upper=0.5):
    """MCNS.py:687-716 with the bounds of mpihans.py:24-33."""
    caps = [50.0, 50.0, None, None, None, None]
    out = dict(steps)
    for i, n in enumerate(STEP_NAMES):
        if move_ratio[i] == 0:
            continue
        if avg_rate[i] < lower:
            out[n] = max(0.5 * steps[n], STEP_MIN[i])
        elif avg_rate[i] > upper:
            out[n] = 2.0 * steps[n] if caps[i] is None else min(2.0 * steps[n], caps[i])
    return out


def run_ns(nchains, nbeads, move_ratio, K, walklength, iterations, seed, initial_walk=50, adjust=True,
           mc_adjust_wl=10):
    """Returns dict(volumes[iterations], steps (final), rates history)."""
    steps = dict(dv_max=2.0, dr_max=2.0, dt_max=0.43, dh_max=0.4, dshear=1.0, dstretch=1.0)  # mpihans.py:118-144
    mr = np.asarray(move_ratio, dtype=float)
    cfg = ho.make_cfg(nchains, nbeads, mr, **steps)
    H, R = ho.grow_walkers(cfg, K, seed)
    ctrs = np.zeros(K, dtype=np.uint64)
    gids = np.arange(K)
    ho.mc_run_many(cfg, H, R, gids, initial_walk, 1e300, seed, gids, ctrs)  # mpihans.py:165
    vols = np.array([ho.volume(H[w]) for w in range(K)])
    rng = np.random.default_rng([seed, 77])
    out = np.zeros(iterations)
    interval = max(K // 2, 1)  # mpihans.py:185
    sH, sR = np.zeros((1, 3, 3)), np.zeros((1,) + R.shape[1:])
    sctr = np.zeros(1, dtype=np.uint64)
    for i in range(iterations):
        imax = int(np.argmax(vols))            # mpihans.py:211-212
        vmax = float(vols[imax])
        out[i] = vmax                          # mpihans.py:221
        src = int(rng.integers(K))             # mpihans.py:220 (may be the max walker itself)
        H[imax], R[imax] = H[src], R[src]      # mpihans.py:226-237
        v, _, _ = ho.mc_run_many(cfg, H, R, [imax], walklength, vmax, seed, [imax], ctrs)  # :241
        vols[imax] = v[0]
        if adjust and i % interval == 0:       # mpihans.py:248, MCNS.py:657-717
            box = int(rng.integers(K))
            rate = np.zeros(6)
            for t in range(6):
                if mr[t] == 0:
                    continue
                one = ho.make_cfg(nchains, nbeads, np.eye(6)[t], **steps)
                sH[0], sR[0] = H[box], R[box]
                _, att, acc = ho.mc_run_many(one, sH, sR, [0], mc_adjust_wl, vmax, seed, [K + 1 + t], sctr)
                rate[t] = acc[0, t] / max(att[0, t], 1)
            steps = update_steps(steps, rate, mr)
            cfg = ho.make_cfg(nchains, nbeads, mr, **steps)
    return dict(volumes=out, steps=steps)
