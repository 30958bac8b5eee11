"""Philox4x32-10 on the host (numpy).  Used where the HOST must agree with draws made on the device
(clone-source selection of the NS loop, tests); the device generator is in csrc/hans_device.cuh."""
from __future__ import annotations

import numpy as np

M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF

STREAM_MOVE, STREAM_GROW, STREAM_NS = 0, 1, 2


def philox4x32_10(ctr, key):
    """ctr: 4 uint32, key: 2 uint32 -> 4 uint32 (Random123 reference vectors in tests/)."""
    c0, c1, c2, c3 = (int(x) & MASK for x in ctr)
    k0, k1 = (int(x) & MASK for x in key)
    for r in range(10):
        if r:
            k0 = (k0 + W0) & MASK
            k1 = (k1 + W1) & MASK
        p0 = M0 * c0
        p1 = M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c3 ^ k1) & MASK, p0 & MASK
    return np.array([c0, c1, c2, c3], dtype=np.uint32)


def ns_draw(seed: int, iteration: int, c2: int, blk: int) -> int:
    """64-bit draw of the NS stream (csrc ns_draw): clone source with (c2, blk) = (0, 0), active
    walker of global virtual rank v with (v, 1)."""
    x = philox4x32_10([iteration & MASK, (iteration >> 32) & MASK, c2 & MASK, (blk & 0xFFFF) | (STREAM_NS << 16)],
                      [seed & MASK, (seed >> 32) & MASK])
    return (int(x[0]) << 32) | int(x[1])


def clone_source(seed: int, iteration: int, nlocal: int, world: int):
    """(rank, local index) of the walker cloned at `iteration` (replaces mpihans.py:220)."""
    g = ns_draw(seed, iteration, 0, 0) % (nlocal * world)
    return divmod(g, nlocal)
