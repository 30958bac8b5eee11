"""The per-walker cell list (BASELINE.json north_star; the link cells the reference bypasses at
/root/reference/NesSa/MCNS.py:603): decisions and trajectories must equal the all-pairs oracle's, bit for bit, in the
dilute regime, in the dense regime (where bounding spheres prune nothing), in thin / sheared cells (axes with fewer than
three bins), with explicit proposals and self-driving, and in the default dispatch, where the warp-per-move and the
cooperative kernel send only the box moves of dense trial configurations through a cell list."""
import numpy as np
import pytest

from common import SHAPES, compress, make_walkers

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

CELLS = 4096  # HANS_TPW_CELLS


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


def densify(eng, oracle, cfg, shape, H, R, k, density, seed, max_rounds=900, shrink=0.985, prep_tpw=128):
    """Walkers at a bead number density.  Small walkers: the oracle's own shrink / relax loop (tests/common.py).  Large
    ones (hundreds of beads: minutes on the CPU): the same procedure on the device with the plain cooperative kernel
    (threads_per_walker = 128, no cell list) -- only a way to MAKE dense start states; the comparison that follows is
    against the oracle."""
    N = R.shape[1]
    if N < 200:
        return compress(cfg, H, R, density, seed)
    nchains, nbeads, mr = SHAPES[shape]
    pool = make_pool(eng, shape, H, R, k, seed=seed + 1000, tpw=prep_tpw)
    relax = pool.cfg_with([0.0] + [float(x) for x in mr[1:4]] + [0.0, 0.0])
    relax.dr_max, relax.dt_max, relax.dh_max = 0.3, 0.3, 0.3
    nw = H.shape[0]
    ids = torch.arange(nw, dtype=torch.int32, device=pool.device)
    for _ in range(max_rounds):
        todo = (N / pool.vol) < density
        if not bool(todo.any()):
            break
        Hb, Rb = pool.H.clone(), pool.R.clone()
        sel = ids[todo]
        pool.change_box(sel, ((shrink - 1.0) * pool.H[todo]).cpu().numpy())
        bad = pool.check_overlap().bool()
        pool.H[bad] = Hb[bad]
        pool.R[bad] = Rb[bad]
        pool.refresh_volumes()
        pool.mc_run(2, counters=False, cfg=relax, count_pairs=False)
    H2, R2 = pool.get_state()
    assert ((N / np.array([oracle.volume(h) for h in H2])) >= density * 0.98).all(), "could not reach the density"
    assert not any(oracle.check_overlap(cfg, H2[w], R2[w]) for w in range(nw))
    return H2, R2


def _replay(eng, oracle, shape, tpw, nwalkers=8, nmoves=400, seed=21, dense=None, mode=0, squash=None, presweeps=2, prep_tpw=128,
            **kw):
    cfg, H, R, k = make_walkers(shape, nwalkers, seed=seed, presweeps=presweeps, **kw)
    if dense:
        H, R = densify(eng, oracle, cfg, shape, H, R, k, dense, seed, prep_tpw=prep_tpw)
    if squash is not None:   # affine squash of cell and beads along one axis would overlap; shear instead (volume kept)
        for w in range(nwalkers):
            oracle.change_box(cfg, H[w], R[w], squash(w, H[w]))
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


@pytest.mark.parametrize("shape,T", [("C1", 0), ("C1", 32), ("C1", 64), ("C2", 0), ("C2", 128), ("C4s", 0), ("C4s", 256),
                                     ("dimer", 32), ("tiny", 0), ("tiny", 64), ("tiny1", 0), ("tiny1", 32), ("p7", 64),
                                     ("t3", 128), ("h15", 0), ("d40", 64), ("d140", 128)])
def test_cells_replay_all_moves(eng, oracle, shape, T):
    nchains, nbeads, _ = SHAPES[shape]
    SHAPES["_tmp"] = (nchains, nbeads, [1, 1, 1, 1 if nbeads >= 4 else 0, 1, 1])
    try:
        props, dec = _replay(eng, oracle, "_tmp", CELLS | T)
    finally:
        del SHAPES["_tmp"]
    assert set(np.unique(props["type"])) >= ({0, 1, 2, 4, 5} if nbeads > 1 else {0, 1, 4, 5})
    assert 0 < dec["accepted"].mean() < 1


@pytest.mark.parametrize("shape,T,density", [("C1", 0, 0.45), ("C1", 64, 0.3), ("C2", 0, 0.7), ("C2", 32, 0.7), ("C4s", 128, 0.4),
                                             ("dimer", 0, 0.5), ("d40", 64, 0.3), ("h15", 128, 0.25), ("p7", 0, 0.3),
                                             ("tiny", 32, 0.5), ("t3", 0, 0.3)])
def test_cells_replay_dense(eng, oracle, shape, T, density):
    """Dense regime: cells a few diameters wide (2-4 bins per axis), most moves rejected by overlap."""
    nchains, nbeads, _ = SHAPES[shape]
    SHAPES["_tmp"] = (nchains, nbeads, [2, 6, 6 if nbeads > 1 else 0, 3 if nbeads >= 4 else 0, 2, 2])
    try:
        props, dec = _replay(eng, oracle, "_tmp", CELLS | T, nwalkers=6, nmoves=500, dense=density, dr_max=0.25, dt_max=0.3,
                             dh_max=0.3, dv_max=2.0, dshear=0.1, dstretch=0.05)
    finally:
        del SHAPES["_tmp"]
    assert (dec["reason"] == 1).mean() > 0.02 and dec["accepted"].mean() > 0.05


@pytest.mark.parametrize("T", [0, 32, 128])
def test_cells_thin_and_sheared_cells(eng, oracle, T):
    """Strongly sheared cells (perpendicular widths down to ~2 sigma: one or two bins along an axis, where the wrapped
    image need not be the nearest one) -- the list must still return every partner the single-image test flags."""
    def shear(w, H):
        d = np.zeros((3, 3))
        d[0] = (0.9 + 0.2 * w) * H[1] + 0.4 * H[2]     # a += 0.9..1.9 b + 0.4 c : volume kept, widths shrink
        d[2] = -0.5 * H[1]
        return d
    props, dec = _replay(eng, oracle, "tiny", CELLS | T, nwalkers=6, nmoves=400, dense=0.35, squash=shear, dr_max=0.5,
                         dshear=0.3, dstretch=0.1, min_ar=0.05, min_ang=3.0)
    assert 0 < dec["accepted"].mean() < 1
    props, dec = _replay(eng, oracle, "tiny1", CELLS | T, nwalkers=6, nmoves=400, dense=0.45, squash=shear, dr_max=0.5,
                         dshear=0.3, dstretch=0.1, min_ar=0.05, min_ang=3.0)
    assert 0 < dec["accepted"].mean() < 1


def test_cells_propose_mode(eng, oracle):
    _replay(eng, oracle, "C1", CELLS | 32, nwalkers=6, nmoves=60, mode=1)
    _replay(eng, oracle, "C4s", CELLS | 128, nwalkers=4, nmoves=60, mode=1, dense=0.4)


@pytest.mark.parametrize("shape,T,sweeps,density", [("C1", 0, 5, None), ("C2", 0, 8, 0.6), ("C4s", 128, 3, 0.4), ("C4", 0, 1, None),
                                                    ("C4", 128, 1, 0.5), ("C5", 0, 1, None), ("C5", 128, 1, 0.3),
                                                    ("C5", 256, 1, 0.3), ("d140", 0, 2, 0.25), ("h15", 64, 5, 0.25)])
def test_cells_mc_run_trajectory_bit_exact(eng, oracle, shape, T, sweeps, density):
    """Self-driving walks with the cell list for every walker, up to the full C4 / C5 walkers, dilute and dense."""
    nw = 8 if shape not in ("C4", "C5") else 5
    cfg, H, R, k = make_walkers(shape, nw, seed=33)
    if density:
        H, R = densify(eng, oracle, cfg, shape, H, R, k, density, 33)
    seed = 515151
    limit = float(np.median([oracle.volume(H[w]) for w in range(nw)]))
    pool = make_pool(eng, shape, H, R, k, seed=seed, tpw=CELLS | T)
    ids = [3, 0, 4, 1]
    att, acc = pool.mc_run(sweeps, ids=ids, volume_limit=limit)
    ctrs = np.zeros(nw, dtype=np.uint64)
    o_vol, o_att, o_acc = oracle.mc_run_many(cfg, H, R, ids, sweeps, limit, seed, np.array(ids), ctrs)
    assert np.array_equal(att.cpu().numpy().astype(np.uint64), o_att)
    assert np.array_equal(acc.cpu().numpy().astype(np.uint64), o_acc)
    assert np.array_equal(pool.ctr.cpu().numpy().astype(np.uint64), ctrs)
    assert_state_equal(pool, H, R)
    assert np.array_equal(pool.vol.cpu().numpy()[ids], o_vol)
    assert not pool.check_overlap().cpu().numpy().any()


@pytest.mark.parametrize("shape,density,ceiling", [("C4", 0.5, True), ("C4", 0.5, False), ("L16", 0.45, True), ("p16", 0.45, True),
                                                   ("C5", 0.3, True), ("d140", 0.3, True)])
def test_default_dispatch_dense_walkers(eng, oracle, shape, density, ceiling):
    """threads_per_walker = 0 on large chain systems compressed until bounding spheres of chains prune nothing in box moves
    (long chains: C4's 12-mers, 16-mers) and on short chains (C5, dimers), with and without a ceiling: the default kernels
    -- warp-per-move and cooperative, whose box moves prune on bead level (lst2 in move_box) -- give the oracle's result."""
    SHAPES["L16"] = (24, 16, [3, 24, 24, 12, 3, 3])    # warp-per-move kernel (>= 24 chains)
    SHAPES["p16"] = (14, 16, [3, 14, 14, 8, 3, 3])     # 224 beads, 14 chains: cooperative kernel
    try:
        nw = 5
        cfg, H, R, k = make_walkers(shape, nw, seed=35)
        H, R = densify(eng, oracle, cfg, shape, H, R, k, density, 35)
        seed = 99
        pool = make_pool(eng, shape, H, R, k, seed=seed, tpw=0)
        limit = float(max(oracle.volume(H[w]) for w in range(nw))) * 1.0001 if ceiling else 1e300
        ctrs = np.zeros(nw, dtype=np.uint64)
        for sweeps in (1, 1):
            att, acc = pool.mc_run(sweeps, volume_limit=limit)
            o_vol, o_att, o_acc = oracle.mc_run_many(cfg, H, R, np.arange(nw), sweeps, limit, seed, np.arange(nw), ctrs)
            assert np.array_equal(att.cpu().numpy().astype(np.uint64), o_att)
            assert np.array_equal(acc.cpu().numpy().astype(np.uint64), o_acc)
        assert np.array_equal(pool.ctr.cpu().numpy().astype(np.uint64), ctrs)
        assert_state_equal(pool, H, R)
        assert np.array_equal(pool.vol.cpu().numpy(), o_vol)
        assert o_att[:, [0, 4, 5]].sum() > 0        # box moves were part of the walk
    finally:
        del SHAPES["L16"], SHAPES["p16"]


@pytest.mark.parametrize("tpw", [0, 1, 8, 24, 28, 32, 64])
def test_thin_cells_every_kernel_family(eng, oracle, tpw):
    """Perpendicular widths between ~1.5 and 2.1 sigma: the float32 pre-filter switches itself off (hans_walk.cuh,
    filter_params) and the chain-level pruning keeps everything; every kernel family must still follow the oracle."""
    def shear(w, H):
        d = np.zeros((3, 3))
        d[0] = (0.9 + 0.2 * w) * H[1] + 0.4 * H[2]
        d[2] = -0.5 * H[1]
        return d
    cfg, H, R, k = make_walkers("tiny", 6, seed=21, dr_max=0.5, dshear=0.3, dstretch=0.1, min_ar=0.05, min_ang=3.0)
    H, R = compress(cfg, H, R, 0.35, 21)
    widths = []
    for w in range(6):
        oracle.change_box(cfg, H[w], R[w], shear(w, H[w]))
        Hi = np.linalg.inv(H[w])
        widths.append(1.0 / np.sqrt((Hi ** 2).sum(axis=0)).max())
    assert min(widths) < 2.0 and max(widths) < 3.5, widths
    try:
        _replay(eng, oracle, "tiny", tpw, nwalkers=6, nmoves=400, dense=0.35, squash=shear, dr_max=0.5, dshear=0.3,
                dstretch=0.1, min_ar=0.05, min_ang=3.0)
    except Exception as e:
        from hans_b200 import _lib
        if isinstance(e, _lib.HansError) and "needs" in str(e):
            pytest.skip(str(e))
        raise


# ---- walkers that do not fit the 227 KB of shared memory: state in global memory + cell list -------------------------
BIG = (512, 6, [2, 6, 6, 3, 2, 2])   # 3 072 beads: ~260 KB of walker state


@pytest.mark.parametrize("tpw,density", [(0, None), (0, 0.11), (CELLS | 256, 0.11), (128, None), (CELLS | 32, None)])
def test_walker_beyond_shared_memory_replay(eng, oracle, tpw, density):
    """3 072 beads per walker: the kernels run with the walker state in a per-CTA region of global memory (namespace hbg)
    and, by default, the per-walker cell list for every overlap test.  Explicit proposals of every move type, at the
    initial density and compressed (1.6 x the bead density; further compression of 3 072 beads takes minutes): decisions, reasons, weights and the post-move state equal the all-pairs oracle's, bit for bit."""
    import ctypes as C
    from hans_b200 import _lib
    SHAPES["big"] = BIG
    try:
        assert _lib.lib.hans_walker_smem_bytes(C.byref(_lib.make_cfg(BIG[0], BIG[1], BIG[2]))) > 227 * 1024
        # (no oracle pre-walk: two sweeps of 5 127 all-pairs moves take minutes on the CPU; dense states are prepared on
        # the device with the cell-list kernel and then checked free of overlaps by the oracle)
        props, dec = _replay(eng, oracle, "big", tpw, nwalkers=2, nmoves=160, seed=41, dense=density, presweeps=0,
                             prep_tpw=CELLS | 256, dr_max=0.6 if density else 2.0, dv_max=60.0, dshear=0.2, dstretch=0.05)
        assert set(np.unique(props["type"])) >= {0, 1, 2, 3, 4, 5} and 0 < dec["accepted"].mean() < 1
    finally:
        del SHAPES["big"]


def test_walker_beyond_shared_memory_walk(eng, oracle):
    """Growth, overlap predicate, change_box and a self-driving walk for the same walker size (all through the
    global-memory state path), against the oracle."""
    SHAPES["big"] = BIG
    try:
        nchains, nbeads, mr = BIG
        nw, seed = 2, 4242
        pool = eng.WalkerPool(nw, nchains, nbeads, mr, seed=seed)
        pool.set_initial_cells()
        pool.grow()
        cfg = oracle.make_cfg(nchains, nbeads, mr)
        H, R = oracle.grow_walkers(cfg, nw, seed)
        assert_state_equal(pool, H, R)
        assert not pool.check_overlap().cpu().numpy().any()
        dH = np.array([[[0.3, 0.1, 0.0], [0.0, -0.2, 0.4], [0.2, 0.0, 0.1]]] * nw)
        pool.change_box([0, 1], dH)
        for w in range(nw):
            oracle.change_box(cfg, H[w], R[w], dH[w])
        assert_state_equal(pool, H, R)
        got = pool.check_overlap().cpu().numpy()
        assert np.array_equal(got, [oracle.check_overlap(cfg, H[w], R[w]) for w in range(nw)])
        # a self-driving walk of 300 moves (hans_mc_run_batch walks whole sweeps of 5 127 moves: replay its proposals)
        nm = 300
        props = np.stack([oracle.gen_proposals(cfg, seed, w, 0, nm) for w in range(nw)])
        limit = float(max(oracle.volume(H[w]) for w in range(nw)))
        got = pool.replay(None, props, volume_limit=limit)
        want = np.stack([oracle.replay(cfg, H[w], R[w], props[w], limit) for w in range(nw)])
        assert np.array_equal(got["accepted"], want["accepted"]) and np.array_equal(got["reason"], want["reason"])
        assert_state_equal(pool, H, R)
    finally:
        del SHAPES["big"]
