"""GPU <-> CPU-oracle parity through the C ABI (libhans_b200.so).

Bar: BIT-EXACT.  Accept/reject decisions, reasons, boltz factors and the float64 post-move state
(cell + every bead) must be identical to the oracle's on the same proposals and configurations
(BASELINE.json north_star, correctness level 1).
"""
import numpy as np
import pytest

from common import SHAPES, compress, make_walkers

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def eng():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from hans_b200 import engine
    return engine


def make_pool(eng, shape, H, R, kw, seed=0, tpw=0):
    nchains, nbeads, mr = SHAPES[shape]
    pool = eng.WalkerPool(H.shape[0], nchains, nbeads, mr, seed=seed, threads_per_walker=tpw, **kw)
    pool.set_state(H, R)
    return pool


def assert_state_equal(pool, H, R):
    gH, gR = pool.get_state()
    assert np.array_equal(gH.view(np.uint64), H.view(np.uint64)), "cell differs from oracle"
    bad = np.nonzero((gR.view(np.uint64) != R.view(np.uint64)).any(axis=(1, 2)))[0]
    assert bad.size == 0, f"bead coordinates differ from oracle for walkers {bad[:8]}"


# ---- a11: overlap predicate --------------------------------------------------------------------
@pytest.mark.parametrize("shape", ["C1", "C2", "C4s", "dimer"])
@pytest.mark.parametrize("tpw", [32, 64, 128, 256])
def test_overlap_predicate(eng, oracle, shape, tpw):
    cfg, H, R, kw = make_walkers(shape, 24, seed=11)
    rng = np.random.default_rng(5)
    # adversarial variants: random displacement of one bead so that roughly half the walkers overlap,
    # wrapped copies (bead moved by a lattice vector), sheared cells
    for w in range(H.shape[0]):
        if w % 3 == 0:
            i, j = rng.choice(R.shape[1], 2, replace=False)
            d = rng.normal(size=3)
            R[w, i] = R[w, j] + d / np.linalg.norm(d) * rng.uniform(0.9, 1.1)
        if w % 4 == 1:
            i = rng.integers(R.shape[1])
            R[w, i] += rng.integers(-2, 3, size=3) @ H[w]
        if w % 5 == 2:
            H[w, 0] += 0.3 * H[w, 1]
    pool = make_pool(eng, shape, H, R, kw, tpw=tpw)
    for inter_only in (False, True):
        got = pool.check_overlap(inter_only=inter_only).cpu().numpy()
        want = np.array([oracle.check_overlap(cfg, H[w], R[w], inter_only) for w in range(H.shape[0])])
        assert np.array_equal(got, want)
        assert 0 < want.sum() < len(want)
    # subset through ids
    ids = [5, 0, 17]
    got = pool.check_overlap(ids=ids).cpu().numpy()
    assert np.array_equal(got, [oracle.check_overlap(cfg, H[w], R[w]) for w in ids])


def test_overlap_near_contact(eng, oracle):
    """Pairs placed within a few ulp of r = sigma: strict r^2 < 1 must agree bit for bit."""
    nchains, nbeads, mr = 2, 1, [1, 6, 0, 0, 3, 3]
    cfg = oracle.make_cfg(nchains, nbeads, mr)
    n = 256
    H = np.tile(np.diag([5.0, 6.0, 7.0]), (n, 1, 1))
    H[:, 1, 0] = 0.7
    R = np.zeros((n, 2, 3))
    rng = np.random.default_rng(3)
    for w in range(n):
        d = rng.normal(size=3)
        d /= np.linalg.norm(d)
        R[w, 0] = rng.uniform(-3, 3, size=3)
        R[w, 1] = R[w, 0] + d * (1.0 + (w - n // 2) * 2.0 ** -52) + rng.integers(-1, 2, size=3) @ H[w]
    from hans_b200 import engine
    pool = engine.WalkerPool(n, nchains, nbeads, mr)
    pool.set_state(H, R)
    got = pool.check_overlap().cpu().numpy()
    want = np.array([oracle.check_overlap(cfg, H[w], R[w]) for w in range(n)])
    assert np.array_equal(got, want)
    assert 0 < want.sum() < n


# ---- a12/a13/a16: cell metrics ------------------------------------------------------------------
def test_cell_metrics(eng, oracle):
    rng = np.random.default_rng(1)
    n = 64
    H = np.tile(np.eye(3) * 7.0, (n, 1, 1)) + rng.normal(scale=1.5, size=(n, 3, 3))
    from hans_b200 import engine
    pool = engine.WalkerPool(n, 4, 2, [1, 1, 1, 0, 1, 1])
    pool.H.copy_(torch.as_tensor(H))
    v, a, c = (t.cpu().numpy() for t in pool.cell_metrics())
    for w in range(n):
        assert v[w] == oracle.volume(H[w])
        assert a[w] == oracle.min_aspect_ratio(H[w])
        assert c[w] == oracle.max_abs_cos(H[w])


# ---- a4..a10, a14, a15: every move type in replay mode --------------------------------------------
def _replay_case(eng, oracle, shape, tpw, nwalkers=12, nmoves=400, seed=21, dense=False, mode=0, **kw):
    cfg, H, R, k = make_walkers(shape, nwalkers, seed=seed, **kw)
    if dense:
        H, R = compress(cfg, H, R, dense, seed)
    props = np.stack([oracle.gen_proposals(cfg, 99, w, 0, nmoves) for w in range(nwalkers)])
    limit = float(np.max([oracle.volume(H[w]) for w in range(nwalkers)]))
    pool = make_pool(eng, shape, H, R, k, tpw=tpw)
    got = pool.replay(None, props, volume_limit=limit, mode=mode)
    want = np.stack([oracle.replay(cfg, H[w], R[w], props[w], limit, mode) for w in range(nwalkers)])
    for f in ("accepted", "reason"):
        assert np.array_equal(got[f], want[f]), f
    assert np.array_equal(got["boltz"].view(np.uint64), want["boltz"].view(np.uint64))
    assert_state_equal(pool, H, R)
    return props, want


# tpw 32..256: cooperative kernel; tpw 1 / 8: warp-per-walker kernel, one move at a time / speculative
# windows of 8 moves; 24 / 28: one CTA of 4 / 8 warps per walker, one warp per move of a window; 0: default
@pytest.mark.parametrize("shape,tpw", [("C1", 32), ("C1", 128), ("C2", 32), ("C2", 64), ("C4s", 32), ("C4s", 256),
                                        ("dimer", 32), ("C1", 0), ("C1", 1), ("C1", 8), ("C2", 0), ("C2", 1), ("C2", 8),
                                        ("C4s", 0), ("C4s", 1), ("C4s", 8), ("dimer", 0), ("dimer", 1), ("dimer", 8),
                                        ("tiny", 0), ("tiny", 8), ("tiny", 32), ("tiny", 64), ("tiny1", 0), ("tiny1", 32),
                                        ("tiny1", 128), ("C1", 24), ("C1", 28), ("C4s", 24), ("C4s", 28), ("dimer", 24),
                                        ("tiny", 28), ("p7", 24), ("t3", 28)])
def test_replay_all_moves(eng, oracle, shape, tpw):
    # equal weights so that every move type (incl. dihedral where nbeads >= 4) is exercised
    nchains, nbeads, _ = SHAPES[shape]
    SHAPES["_tmp"] = (nchains, nbeads, [1, 1, 1, 1 if nbeads >= 4 else 0, 1, 1])
    try:
        props, dec = _replay_case(eng, oracle, "_tmp", tpw)
    finally:
        del SHAPES["_tmp"]
    types = set(np.unique(props["type"]))
    assert types >= ({0, 1, 2, 4, 5} if nbeads > 1 else {0, 1, 4, 5})
    # both outcomes seen
    assert 0 < dec["accepted"].mean() < 1


@pytest.mark.parametrize("tpw", [32, 1, 8])
def test_replay_dense_rejections(eng, oracle, tpw):
    """Dense regime: most moves are rejected by overlap, early exit paths are hit."""
    props, dec = _replay_case(eng, oracle, "C2", tpw, nwalkers=8, nmoves=600, dense=0.7, dr_max=0.6, dv_max=3.0)
    assert (dec["reason"] == 1).mean() > 0.2


@pytest.mark.parametrize("shape,tpw,density", [("d40", 0, 0.3), ("t3", 0, 0.3), ("h15", 0, 0.2), ("p7", 64, 0.3), ("C1", 0, 0.15), ("C1", 64, 0.15), ("C1", 0, 0.45), ("C1", 32, 0.45), ("C1", 128, 0.3), ("C1", 8, 0.45),
                                                ("C4s", 0, 0.4), ("C4s", 64, 0.4), ("dimer", 0, 0.5), ("tiny", 0, 0.5),
                                                ("C1", 24, 0.45), ("C1", 28, 0.3), ("C4s", 28, 0.4), ("d40", 24, 0.3),
                                                ("h15", 28, 0.2), ("p7", 24, 0.3), ("tiny", 28, 0.5)])
def test_replay_dense_chains(eng, oracle, shape, tpw, density):
    """Chain systems compressed until the chains' bounding spheres touch many neighbours and the cell is only a
    few diameters wide: the chain-level pruning keeps most neighbours (or, once the bounding distance exceeds half
    the cell width, all of them) and decisions must still be the checker's, bit for bit."""
    nchains, nbeads, _ = SHAPES[shape]
    SHAPES["_tmp"] = (nchains, nbeads, [2, 6, 6, 3 if nbeads >= 4 else 0, 2, 2])
    try:
        props, dec = _replay_case(eng, oracle, "_tmp", tpw, nwalkers=6, nmoves=500, dense=density, dr_max=0.25, dt_max=0.3,
                                  dh_max=0.3, dv_max=2.0, dshear=0.1, dstretch=0.05)
    finally:
        del SHAPES["_tmp"]
    assert (dec["reason"] == 1).mean() > 0.02 and dec["accepted"].mean() > 0.05


@pytest.mark.parametrize("shape,tpw", [("C2", 8), ("C1", 8), ("C4s", 8), ("dimer", 8), ("C1", 24), ("C1", 28), ("C4s", 28),
                                        ("dimer", 24), ("d40", 28), ("h15", 24)])
def test_replay_windows_only_chain_moves(eng, oracle, shape, tpw):
    """No box moves at all: every window of the speculative path is full, and with small steps most
    moves are accepted, so the in-window fix-ups and stale re-evaluations carry the result."""
    nchains, nbeads, _ = SHAPES[shape]
    SHAPES["_tmp"] = (nchains, nbeads, [0, 3, 2 if nbeads > 1 else 0, 1 if nbeads >= 4 else 0, 0, 0])
    try:
        props, dec = _replay_case(eng, oracle, "_tmp", tpw, nwalkers=8, nmoves=700, dr_max=0.3, dt_max=0.2, dh_max=0.2)
    finally:
        del SHAPES["_tmp"]
    assert dec["accepted"].mean() > 0.5


@pytest.mark.parametrize("tpw", [0, 1, 8, 24, 28, 64])
def test_replay_band_hits(eng, oracle, tpw):
    """Contacts at 1 +- 1e-9 .. 1e-13: the float32 filter cannot decide (its band is ~1e-5 wide), float64 must,
    in every kernel -- in the windowed kernels also when the partner is a chain an EARLIER move of the same window
    has just put there (proposals a_i, b_i below)."""
    from hans_b200 import _lib
    nw, L = 4, 20.0
    nchains, nbeads, mr = SHAPES["dimer"]
    k = dict(dv_max=2.0, dr_max=2.0, dt_max=0.43, dh_max=0.4, dshear=1.0, dstretch=1.0)
    cfg = oracle.make_cfg(nchains, nbeads, mr, **k)
    H = np.tile(np.diag([L, L, L]), (nw, 1, 1)).astype(np.float64)
    sites = np.array([(1.5 + 3.0 * i, 1.5 + 3.0 * j, 3.0) for i in range(6) for j in range(6)])[:nchains]
    R = np.zeros((nw, nchains * nbeads, 3))
    R[:, 0::2] = sites
    R[:, 1::2] = sites + np.array([0.4, 0.0, 0.0])
    deltas = [1e-9, -1e-9, 1e-12, -1e-12, 1e-13, -1e-13, 3e-7, -3e-7, 0.0, 2e-16]
    props = np.zeros((nw, 22), dtype=oracle.PROPOSAL_DTYPE)
    props["type"] = 1
    props["u_acc"] = 0.5
    for w in range(nw):
        for i in range(10):      # mover 10 + i lands on top of anchor i, at 1 + delta
            d = deltas[(i + w) % len(deltas)]
            props["ichain"][w, i] = 10 + i
            props["p"][w, i, :3] = sites[i] + np.array([0.0, 0.0, 1.0 + d]) - sites[10 + i]
        for i in range(6):       # a_i = chain 20 + i goes up; b_i = chain 26 + i lands beside its NEW position
            d = deltas[(3 * i + w) % len(deltas)]
            props["ichain"][w, 10 + 2 * i] = 20 + i
            props["p"][w, 10 + 2 * i, :3] = [0.0, 0.0, 5.0]
            props["ichain"][w, 11 + 2 * i] = 26 + i
            props["p"][w, 11 + 2 * i, :3] = sites[20 + i] + np.array([0.0, 1.0 + d, 5.0]) - sites[26 + i]
    H0, R0 = H.copy(), R.copy()
    want = np.stack([oracle.replay(cfg, H[w], R[w], props[w], 1e300, 0) for w in range(nw)])  # (advances H, R)
    assert 0.2 < want["accepted"].mean() < 0.95
    _lib.diag_counters(reset=True)
    pool = make_pool(eng, "dimer", H0, R0, k, tpw=tpw)
    got = pool.replay(None, props, volume_limit=1e300, mode=0)
    for f in ("accepted", "reason"):
        assert np.array_equal(got[f], want[f]), f
    assert_state_equal(pool, H, R)
    assert _lib.diag_counters()[0] >= nw * 10  # the float64 settle really ran


@pytest.mark.parametrize("shape,tpw", [("dimer", 0), ("dimer", 8), ("dimer", 32), ("C2", 0), ("C2", 8), ("C1", 0), ("C1", 24),
                                       ("C4s", 28), ("C4s", 128)])
def test_replay_u_acc_at_or_above_one_is_a_rejection(eng, oracle, shape, tpw):
    """MCNS.py:372 accepts a single-chain move iff U < boltz, boltz = 1 without overlap.  Generated U lie in [0, 1); an
    EXPLICIT proposal may carry U >= 1: the move is then rejected (not accepted-but-uncommitted) in every kernel family,
    as in the oracle."""
    nw, nmoves = 5, 48
    cfg, H, R, k = make_walkers(shape, nw, seed=41)
    props = np.stack([oracle.gen_proposals(cfg, 4242, w, 0, nmoves) for w in range(nw)])
    props["u_acc"][:, ::3] = 1.0
    props["u_acc"][:, 1::6] = 1.5
    H0, R0 = H.copy(), R.copy()
    want = np.stack([oracle.replay(cfg, H[w], R[w], props[w], 1e300, 0) for w in range(nw)])  # (advances H, R)
    chain_moves = (props["type"] >= 1) & (props["type"] <= 3)
    assert not want["accepted"][chain_moves & (props["u_acc"] >= 1.0)].any()
    assert want["accepted"][chain_moves & (props["u_acc"] < 1.0)].any()
    pool = make_pool(eng, shape, H0, R0, k, tpw=tpw)
    got = pool.replay(None, props, volume_limit=1e300, mode=0)
    for f in ("accepted", "reason"):
        assert np.array_equal(got[f], want[f]), f
    assert_state_equal(pool, H, R)


def test_replay_propose_mode(eng, oracle):
    """hs_alkane entry-point semantics: trial state left in place, boltz returned."""
    _replay_case(eng, oracle, "C1", 32, nwalkers=6, nmoves=60, mode=1)


def test_replay_invalid_and_edge(eng, oracle):
    """dihedral on short chains, dV driving the volume negative, u_acc = 0, ids subset."""
    cfg, H, R, k = make_walkers("dimer", 4, seed=2)
    props = np.zeros((4, 4), dtype=oracle.PROPOSAL_DTYPE)
    props["type"][:, 0] = 3; props["idx0"][:, 0] = 3          # dihedral impossible on dimers
    props["type"][:, 1] = 0; props["p"][:, 1, 0] = -1e9        # V + dV <= 0
    props["type"][:, 2] = 0; props["p"][:, 2, 0] = -1.0; props["u_acc"][:, 2] = 0.0
    props["type"][:, 3] = 1; props["p"][:, 3, :3] = [100.0, -250.0, 1e-9]; props["ichain"][:, 3] = 31
    pool = make_pool(eng, "dimer", H, R, k)
    got = pool.replay([2, 0], props[:2], volume_limit=1e300)
    want = np.stack([oracle.replay(cfg, H[w], R[w], props[0], 1e300) for w in (2, 0)])
    assert np.array_equal(got["accepted"], want["accepted"]) and np.array_equal(got["reason"], want["reason"])
    assert list(want["reason"][0][:2]) == [6, 6]
    assert_state_equal(pool, H, R)


# ---- a1: the fused MC_run, self-driving from Philox ------------------------------------------------
@pytest.mark.parametrize("shape,tpw,sweeps", [("C1", 32, 6), ("C1", 64, 3), ("C2", 32, 10), ("C2", 128, 4),
                                               ("C4s", 32, 3), ("C4s", 256, 2), ("dimer", 32, 5),
                                               ("C1", 0, 6), ("C1", 1, 6), ("C1", 8, 4), ("C2", 0, 10), ("C2", 8, 6),
                                               ("C2", 1, 6), ("C4s", 0, 3), ("C4s", 8, 3), ("dimer", 0, 5), ("dimer", 8, 5),
                                               ("tiny", 0, 12), ("tiny", 32, 12), ("tiny1", 0, 12), ("tiny1", 64, 12),
                                               ("C4", 0, 1), ("C4", 128, 1), ("C5", 0, 1), ("C5", 128, 1), ("C5", 256, 1),
                                               ("d40", 0, 6), ("d40", 32, 6), ("d40", 64, 6), ("t3", 0, 8), ("t3", 64, 8),
                                               ("h15", 0, 6), ("h15", 8, 6), ("h15", 128, 6), ("p7", 0, 6), ("p7", 64, 6),
                                               ("C1", 24, 6), ("C4s", 28, 3), ("C4", 24, 1), ("C4", 28, 1), ("C5", 24, 1),
                                               ("C5", 28, 1), ("d40", 28, 6), ("h15", 24, 6), ("p7", 28, 6), ("t3", 24, 8),
                                               ("d140", 0, 2), ("d140", 24, 2), ("d140", 128, 2), ("d140", 32, 2)])
def test_mc_run_trajectory_bit_exact(eng, oracle, shape, tpw, sweeps):
    nw = 10 if shape not in ("C4", "C5") else 7
    cfg, H, R, k = make_walkers(shape, nw, seed=31)
    seed = 424242
    limit = float(np.median([oracle.volume(H[w]) for w in range(nw)]))  # ceiling bites for half
    pool = make_pool(eng, shape, H, R, k, seed=seed, tpw=tpw)
    pool.gid_base = 1000
    ids = [7, 2, 9, 0, 4] if nw == 10 else [5, 2, 6, 0]
    att, acc = pool.mc_run(sweeps, ids=ids, volume_limit=limit)
    att2, acc2 = pool.mc_run(1, ids=ids, volume_limit=limit)  # continues the per-walker counters
    ctrs = np.zeros(nw, dtype=np.uint64)
    o_vol, o_att, o_acc = oracle.mc_run_many(cfg, H, R, ids, sweeps, limit, seed, 1000 + np.array(ids), ctrs)
    o_vol, o_att2, o_acc2 = oracle.mc_run_many(cfg, H, R, ids, 1, limit, seed, 1000 + np.array(ids), ctrs)
    assert np.array_equal(att.cpu().numpy().astype(np.uint64), o_att)
    assert np.array_equal(acc.cpu().numpy().astype(np.uint64), o_acc)
    assert np.array_equal(acc2.cpu().numpy().astype(np.uint64), o_acc2)
    assert np.array_equal(pool.ctr.cpu().numpy().astype(np.uint64), ctrs)
    assert_state_equal(pool, H, R)
    assert np.array_equal(pool.vol.cpu().numpy()[ids], o_vol)
    nmoves = (sweeps + 1) * oracle.moves_per_sweep(cfg.nbeads, cfg.nchains)
    assert int(att.sum() + att2.sum()) == nmoves * len(ids)
    assert int(pool.pair_tests.item()) > 0
    assert not pool.check_overlap().cpu().numpy().any()


@pytest.mark.parametrize("shape,nw", [("C1", 1700), ("C2", 1800), ("dimer", 1700), ("C1", 2100)])
def test_large_batch_uses_the_register_capped_kernel(eng, oracle, shape, nw):
    """More walkers than the uncapped warp-per-walker kernel keeps resident (9 per SM): the launcher switches to a
    register-capped build of the same kernel -- 168 registers for chains at 10 to 12 walkers per SM (1 700 on 148 SMs),
    128 registers for single spheres -- and back to the uncapped one for chains beyond that (2 100).  Same trajectories,
    bit for bit."""
    sweeps, seed = 1, 31337
    cfg, H, R, k = make_walkers(shape, nw, seed=17, presweeps=0)
    pool = make_pool(eng, shape, H, R, k, seed=seed)
    limit = float(np.median([oracle.volume(H[w]) for w in range(0, nw, 50)]))
    att, acc = pool.mc_run(sweeps, volume_limit=limit)
    ctrs = np.zeros(nw, dtype=np.uint64)
    o_vol, o_att, o_acc = oracle.mc_run_many(cfg, H, R, np.arange(nw), sweeps, limit, seed, np.arange(nw), ctrs)
    assert np.array_equal(acc.cpu().numpy().astype(np.uint64), o_acc)
    assert np.array_equal(pool.ctr.cpu().numpy().astype(np.uint64), ctrs)
    assert_state_equal(pool, H, R)
    assert np.array_equal(pool.vol.cpu().numpy(), o_vol)


@pytest.mark.parametrize("shape,tpw,ws_moves", [("C2", 0, 0), ("C2", 0, 96), ("C1", 0, 64), ("C1", 32, 32), ("C4s", 128, 0),
                                                 ("C4s", 128, 160), ("dimer", 1, 32), ("C4s", 24, 160), ("C1", 28, 64)])
def test_mc_run_proposals_ahead(eng, oracle, shape, tpw, ws_moves):
    """hans_mc_run_batch_ws: proposals generated ahead of the walk into a workspace (default), with a
    workspace so small that the walk is cut into chunks, and without one (generation in the kernel)."""
    nw, sweeps, seed = 6, 4, 777
    cfg, H, R, k = make_walkers(shape, nw, seed=13)
    pool = make_pool(eng, shape, H, R, k, seed=seed, tpw=tpw)
    pool.workspace_bytes = ws_moves * nw * oracle.PROPOSAL_DTYPE.itemsize
    limit = float(np.median([oracle.volume(H[w]) for w in range(nw)]))
    att, acc = pool.mc_run(sweeps, volume_limit=limit)
    ctrs = np.zeros(nw, dtype=np.uint64)
    o_vol, o_att, o_acc = oracle.mc_run_many(cfg, H, R, np.arange(nw), sweeps, limit, seed, np.arange(nw), ctrs)
    assert np.array_equal(att.cpu().numpy().astype(np.uint64), o_att)
    assert np.array_equal(acc.cpu().numpy().astype(np.uint64), o_acc)
    assert np.array_equal(pool.ctr.cpu().numpy().astype(np.uint64), ctrs)
    assert_state_equal(pool, H, R)
    assert np.array_equal(pool.vol.cpu().numpy(), o_vol)


def test_gen_proposals_match_oracle(eng, oracle):
    """hans_gen_proposals writes exactly the proposals the checker draws."""
    import ctypes as C
    from hans_b200 import _lib
    nchains, nbeads, mr = SHAPES["C1"]
    pool = eng.WalkerPool(3, nchains, nbeads, [1, 1, 1, 1, 1, 1], seed=99, gid_base=5)
    pool.ctr += torch.tensor([0, 17, 1 << 33], device=pool.device)
    nm = 200
    buf = torch.zeros(2 * nm * oracle.PROPOSAL_DTYPE.itemsize, dtype=torch.uint8, device=pool.device)
    ids = torch.tensor([2, 1], dtype=torch.int32, device=pool.device)
    _lib.check(_lib.lib.hans_gen_proposals(C.byref(pool.cfg), pool.ctr.data_ptr(), ids.data_ptr(), 2, nm, 99, 5,
                                           buf.data_ptr(), None))
    got = buf.cpu().numpy().view(oracle.PROPOSAL_DTYPE).reshape(2, nm)
    cfg = oracle.make_cfg(nchains, nbeads, [1, 1, 1, 1, 1, 1])
    for row, (w, c0) in enumerate([(2, 1 << 33), (1, 17)]):
        want = oracle.gen_proposals(cfg, 99, 5 + w, c0, nm)
        for f in ("type", "ichain", "idx0", "idx1"):
            assert np.array_equal(got[row][f], want[f]), f
        assert np.array_equal(got[row]["p"].view(np.uint64), want["p"].view(np.uint64))
        assert np.array_equal(got[row]["u_acc"].view(np.uint64), want["u_acc"].view(np.uint64))


def test_mc_run_device_limit_and_no_counters(eng, oracle):
    cfg, H, R, k = make_walkers("C1", 4, seed=8)
    pool = make_pool(eng, "C1", H, R, k, seed=5)
    lim = torch.tensor([1400.0], dtype=torch.float64, device=pool.device)
    a, b = pool.mc_run(2, d_volume_limit=lim, counters=False)
    assert a is None and b is None
    ctrs = np.zeros(4, dtype=np.uint64)
    oracle.mc_run_many(cfg, H, R, np.arange(4), 2, 1400.0, 5, np.arange(4), ctrs)
    assert_state_equal(pool, H, R)


# ---- a21: initial configurations -----------------------------------------------------------------
@pytest.mark.parametrize("shape", ["C1", "C2", "C4s"])
def test_grow_matches_oracle(eng, oracle, shape):
    nchains, nbeads, mr = SHAPES[shape]
    nw = 6
    pool = eng.WalkerPool(nw, nchains, nbeads, mr, seed=77, gid_base=3)
    pool.set_initial_cells()
    st = pool.grow()
    cfg = oracle.make_cfg(nchains, nbeads, mr)
    H, R = oracle.grow_walkers(cfg, nw, seed=77, gid0=3)
    assert_state_equal(pool, H, R)
    assert (st >= nchains).all()
    assert not pool.check_overlap().cpu().numpy().any()
    v = pool.vol.cpu().numpy()
    assert np.allclose(v, 0.999 ** 3 * 15 * nchains * nbeads)


# ---- a17/a18: NS iteration helpers ---------------------------------------------------------------
def test_argmax_and_clone(eng, oracle):
    cfg, H, R, k = make_walkers("C2", 40, seed=4)
    pool = make_pool(eng, "C2", H, R, k)
    vols = pool.vol.cpu().numpy().copy()
    # introduce a tie: the lowest index must win (list.index semantics, mpihans.py:212)
    H[33] = H[int(np.argmax(vols))]
    H[12] = H[33]
    pool.set_state(H, R)
    vols = pool.vol.cpu().numpy()
    out = pool.argmax().cpu().numpy()
    assert out[0] == vols.max() and int(out[1]) == int(np.argmax(vols))
    pool.clone(src=3, dst=int(out[1]))
    gH, gR = pool.get_state()
    assert np.array_equal(gH[int(out[1])], H[3]) and np.array_equal(gR[int(out[1])], R[3])
    assert pool.vol.cpu().numpy()[int(out[1])] == vols[3]
    # device-side indices + enable flag
    d_src = torch.tensor([5], dtype=torch.int32, device=pool.device)
    d_dst = torch.tensor([6], dtype=torch.int32, device=pool.device)
    off = torch.tensor([0], dtype=torch.int32, device=pool.device)
    pool.clone(0, 0, d_src=d_src, d_dst=d_dst, d_enable=off)
    assert np.array_equal(pool.get_state()[1][6], R[6])
    pool.clone(0, 0, d_src=d_src, d_dst=d_dst)
    assert np.array_equal(pool.get_state()[1][6], R[5])


def test_change_box(eng, oracle):
    cfg, H, R, k = make_walkers("C1", 5, seed=6)
    pool = make_pool(eng, "C1", H, R, k)
    rng = np.random.default_rng(0)
    dH = rng.normal(scale=0.2, size=(3, 3, 3))
    ids = [4, 1, 2]
    pool.change_box(ids, dH)
    for n, w in enumerate(ids):
        oracle.change_box(cfg, H[w], R[w], dH[n])
    assert_state_equal(pool, H, R)
