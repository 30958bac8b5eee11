"""Shared helpers for the parity tests: seeded configurations built with the CPU oracle."""
import numpy as np

from oracle import oracle as ho

# (nchains, nbeads, move_ratio) of the BASELINE.json configs, reduced walker counts
SHAPES = {
    "C1": (16, 6, [3, 48, 32, 0, 3, 3]),
    "C2": (64, 1, [1, 192, 0, 0, 3, 3]),
    "C4s": (8, 12, [3, 32, 32, 16, 3, 3]),   # C4's chain length, fewer chains
    "dimer": (32, 2, [1, 96, 64, 0, 3, 3]),  # the committed input.txt shape
    "C4": (32, 12, [3, 32, 32, 16, 3, 3]),   # BASELINE configs[3], full walker size (384 beads)
    "C5": (128, 6, [1, 384, 256, 384, 3, 3]),  # BASELINE configs[4], full walker size (768 beads)
    "d40": (40, 2, [2, 40, 30, 0, 3, 3]),      # more than 32 chains of more than one bead: two waves of chain spheres
    "t3": (11, 3, [2, 20, 20, 0, 3, 3]),       # nbeads the warp-per-walker kernel is not instantiated for, odd nchains
    "h15": (15, 6, [3, 48, 32, 10, 3, 3]),     # odd number of chains (round robin over chain pairs), dihedrals on
    "p7": (9, 7, [2, 20, 20, 10, 3, 3]),       # nbeads = 7: cooperative kernel, runtime nbeads, intra-chain pairs
    "d140": (140, 2, [2, 60, 40, 0, 3, 3]),    # more than 128 chains: two chunks of chain spheres in the warp-per-move kernel
    "tiny": (6, 2, [1, 18, 12, 0, 3, 3]),    # fewer beads than lanes: several lanes share a bead row in box moves
    "tiny1": (5, 1, [1, 15, 0, 0, 3, 3]),
}


def cfg_kw(dv_max=2.0, dr_max=2.0, dt_max=0.43, dh_max=0.4, dshear=1.0, dstretch=1.0, **kw):
    d = dict(dv_max=dv_max, dr_max=dr_max, dt_max=dt_max, dh_max=dh_max, dshear=dshear, dstretch=dstretch)
    d.update(kw)
    return d


def make_walkers(shape, nwalkers, seed, presweeps=2, **kw):
    """Grown + briefly walked walkers of a named shape; returns (cfg_oracle, H, R, kwargs)."""
    nchains, nbeads, mr = SHAPES[shape]
    k = cfg_kw(**kw)
    cfg = ho.make_cfg(nchains, nbeads, mr, **k)
    H, R = ho.grow_walkers(cfg, nwalkers, seed)
    if presweeps:
        ctrs = np.zeros(nwalkers, dtype=np.uint64)
        ho.mc_run_many(cfg, H, R, np.arange(nwalkers), presweeps, 1e300, seed + 1, np.arange(nwalkers), ctrs)
    return cfg, H, R, k


def compress(cfg, H, R, target_density, seed, max_rounds=4000, shrink=0.985):
    """Bring walkers to a bead number density (dense regime): isotropic shrink steps through the
    oracle's change_box, each kept only if no overlap appears, with short walks in between."""
    nw = H.shape[0]
    N = R.shape[1]
    ctrs = np.full(nw, 1 << 20, dtype=np.uint64)
    for w in range(nw):
        for _ in range(max_rounds):
            if N / ho.volume(H[w]) >= target_density:
                break
            Hb, Rb = H[w].copy(), R[w].copy()
            ho.change_box(cfg, H[w], R[w], (shrink - 1.0) * H[w])
            if ho.check_overlap(cfg, H[w], R[w]):
                H[w], R[w] = Hb, Rb
                ho.mc_run_many(cfg, H, R, [w], 2, ho.volume(H[w]), seed, [w], ctrs)
    return H, R
