"""End-to-end checks of the host layer on a GPU: the alk / MCNS drop-in surface and the mpihans
driver with its file contract (volumes.txt, traj.extxyz, restart)."""
import io
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


ARGS = dict(nwalkers=4, nchains=8, nbeads=4, bondlength=0.4, bondangle=109.7, min_angle=45.0, min_aspect_ratio=0.8)


def _setup_alk(seed=3):
    from hans_b200 import MCNS as NS
    NS.alk.random_set_random_seed(seed)
    np.random.seed(seed)
    NS.initialise_sim_cells(ARGS, quiet=1)
    NS.create_initial_configs(ARGS)
    return NS


def test_alk_surface_and_views(oracle):
    NS = _setup_alk()
    alk = NS.alk
    assert alk.alkane_get_nchains() == 8 and alk.alkane_get_nbeads() == 4
    cell = alk.box_get_cell(1)
    assert cell.shape == (3, 3) and np.allclose(cell, 0.999 * np.eye(3) * np.cbrt(15 * 32))
    assert alk.box_compute_volume(2) == pytest.approx(0.999 ** 3 * 15 * 32)
    for ibox in range(1, 5):
        assert alk.alkane_check_chain_overlap(ibox) == 0
    # the chain getter is a writable view of the box: writes are seen by the next native call
    ch = alk.alkane_get_chain(3, 2)
    assert ch.shape == (4, 3) and ch.flags.writeable
    other = alk.alkane_get_chain(5, 2)
    keep = ch.copy()
    ch[:] = other + 0.01
    assert alk.alkane_check_chain_overlap(2) == 1
    ch[:] = keep
    assert alk.alkane_check_chain_overlap(2) == 0
    # box predicates agree with the Python statements of the reference (MCNS.py:97-133)
    alk.alkane_change_box(1, np.array([[0.3, 0.1, 0.0], [0.0, -0.2, 0.4], [0.2, 0.0, 0.1]]))
    assert alk.box_min_aspect_ratio(1) == pytest.approx(NS.min_aspect_ratio(1), rel=1e-13)
    assert NS.min_angle(1) == pytest.approx(np.arccos(oracle.max_abs_cos(alk.box_get_cell(1))), rel=1e-12)
    assert alk.alkane_check_chain_overlap(1) in (0, 1)
    # bond geometry survives every move type
    for _ in range(20):
        b = alk.alkane_translate_chain(2, 3); assert b in (0.0, 1.0)
        b, q = alk.alkane_rotate_chain(2, 3, 0); assert q.shape == (4,) and abs(q @ q - 1) < 1e-12
        b, bead, ang = alk.alkane_bond_rotate(2, 3, 1); assert 4 <= bead <= 4
    x = alk.alkane_get_chain(2, 3)
    assert np.allclose(np.linalg.norm(np.diff(x, axis=0), axis=1), 0.4, atol=1e-12)
    # volume move: save / propose / reset restores exactly
    H0, R0 = alk.box_get_cell(4).copy(), alk.alkane_get_chain(1, 4).copy()
    boltz = alk.alkane_box_resize(0.0, 4, 0)
    assert boltz >= 0.0 and not np.array_equal(alk.box_get_cell(4), H0)
    alk.alkane_box_resize(0.0, 4, 1)
    assert np.array_equal(alk.box_get_cell(4), H0) and np.array_equal(alk.alkane_get_chain(1, 4), R0)
    alk.alkane_destroy(); alk.box_destroy()


def test_mc_run_dropin_equals_oracle_walk(oracle):
    """hans_b200.MCNS.MC_run -- the documented drop-in for MCNS.py:276-392 -- returns what the CPU oracle's MC_run returns
    for the same box, seed and counters: volume, the six acceptance rates and the state, bit for bit.  (The reference's own
    per-move loop is exercised on the alk entry points by test_gpu_reference_golden.py.)"""
    NS = _setup_alk(seed=5)
    alk = NS.alk
    mr = NS.default_move_ratio(dict(ARGS))
    vol0 = alk.box_compute_volume(1)
    H0, R0 = alk.box_get_cell(1).copy(), alk._S.R[0].copy()
    v, rate_f = NS.MC_run(ARGS, 60, mr, 1, volume_limit=vol0, min_ang=45.0, min_ar=0.8)
    assert v <= vol0 * (1 + 1e-12) and rate_f.shape == (6,)
    assert alk.alkane_check_chain_overlap(1) == 0
    cfg = oracle.make_cfg(ARGS["nchains"], ARGS["nbeads"], mr, min_ar=0.8, min_ang=45.0)
    ov, att, acc, _ = oracle.mc_run(cfg, H0, R0, 60, vol0, 5, 0, 0)
    assert v == ov and np.array_equal(rate_f, acc / np.where(att == 0, 1, att))
    assert np.array_equal(alk.box_get_cell(1), H0) and np.array_equal(alk._S.R[0], R0)
    at = NS.MC_run(ARGS, 1, mr, 1, volume_limit=vol0, min_ang=45.0, min_ar=0.8, return_ase=True)
    assert len(at) == 32 and np.array_equal(at.positions, alk._S.R[0])
    vols = NS.perturb_initial_configs(dict(ARGS), mr, 5)
    assert set(vols) == {1, 2, 3, 4}
    r, dsh, dst = NS.adjust_mc_steps(dict(ARGS), NS.SingleComm(), mr, vol0, walklength=4,
                                     min_dstep=np.full(6, 1e-10), dv_max=50.0, dr_max=50.0)
    assert r.shape == (6,) and dsh > 0 and dst > 0
    # one alkane_grow_chain call per attempt, the way MCNS.py:634-644 drives it
    for ichain in range(1, ARGS["nchains"] + 1):
        alk.alkane_set_nchains(ichain)
        while True:
            rb, ifail = alk.alkane_grow_chain(ichain, 1, 1)
            if rb != 0 and ifail == 0:
                break
    assert alk.alkane_check_chain_overlap(1) == 0
    # walker copy helpers (MCNS.py:30-53,510-524,550-580)
    NS.clone_walker(1, 3)
    assert np.array_equal(alk._S.R[2], alk._S.R[0]) and np.array_equal(alk.box_get_cell(3), alk.box_get_cell(1))
    at = NS.mk_ase_config(1, ARGS["nbeads"], ARGS["nchains"], scaling=2.0)
    NS.import_ase_to_ibox(at, 2, dict(ARGS), scaling=0.5)
    assert np.array_equal(alk._S.R[1], alk._S.R[0]) and np.array_equal(alk.box_get_cell(2), alk.box_get_cell(1))
    alk.alkane_destroy()


def test_driver_files_and_restart(tmp_path, monkeypatch):
    from hans_b200 import NSio, mpihans
    monkeypatch.chdir(tmp_path)
    text = """move_ratio = 3,48,32,16,3,3
nbeads = 4
nchains = 8
nwalkers = 6
ranks_per_gpu = 4
iterations = 230
walklength = 3
initial_walk = 10
directory = run1
seed = 11
restart_interval = 200
"""
    P = NSio.parse_hans_text(text)
    mpihans.main(dict(P))
    os.chdir(tmp_path)
    lines = open("run1/volumes.txt").read().splitlines()
    assert lines[0] == "24 1 32 False 8 "           # K, n_cull, dof = 8 + 3*8, flat_V_prior, nchains
    assert len(lines) == 231
    it = [int(l.split()[0]) for l in lines[1:]]
    vol = np.array([float(l.split()[1]) for l in lines[1:]])
    assert it == list(range(230)) and (np.diff(vol) <= 1e-9).all(), "ceiling must decrease monotonically"
    assert all(l.endswith(" ") and l.split()[1] == l.split()[2] for l in lines[1:])
    from hans_b200.atoms import read_extxyz
    frames = read_extxyz("run1/traj.extxyz", ":")
    assert len(frames) == 3 and len(frames[0]) == 32  # iterations 0, 100, 200
    frac = np.linalg.solve(frames[1].cell.T, frames[1].positions.T).T
    assert (frac > -1e-6).all() and (frac < 1 + 1e-6).all()
    assert abs(frames[1].get_volume() - vol[100]) < 1e-6 * vol[100]
    attrs, walkers = NSio.open_restart("run1/restart.hdf5")
    assert attrs["prev_iters"] == 230 and len(walkers) == 24 and attrs["dv_max"] > 0
    assert NSio.restart_exists("run1/restart_backup.hdf5")
    # groups are named per VIRTUAL rank, nwalkers each: what 4 MPI ranks x 6 walkers of the reference would write
    # (NSio.py:87, mpihans.py:171-179), independent of ranks_per_gpu
    assert sorted(walkers) == [f"walker_{r}_{i:04d}" for r in range(4) for i in range(1, 7)]
    assert walkers["walker_0_0001"][0].shape == (32, 3) and walkers["walker_3_0006"][1].shape == (3, 3)
    assert all(w[2] is not None and w[2] > 0 for w in walkers.values())   # Philox counters travel with the walkers
    # resume: `iterations` is restored from the file (as in the reference), so 230 more run
    P2 = NSio.parse_hans_text("from_restart = 1\ndirectory = run1\n")
    mpihans.main(dict(P2))
    os.chdir(tmp_path)
    lines = open("run1/volumes.txt").read().splitlines()
    assert len(lines) == 1 + 460 and int(lines[-1].split()[0]) == 459
    vol = np.array([float(l.split()[1]) for l in lines[1:]])
    assert (np.diff(vol) <= 1e-9).all()


def test_resumed_run_equals_uninterrupted_run(tmp_path, monkeypatch):
    """230 + 230 iterations through a restart file write the same volumes.txt, line for line, as 460 iterations in one go
    (same seed): the restart carries every walker's Philox counter and the step adaptation is a pure function of the
    iteration number."""
    from hans_b200 import NSio, mpihans
    monkeypatch.chdir(tmp_path)
    base = """move_ratio = 3,48,32,16,3,3
nbeads = 4
nchains = 8
nwalkers = 6
ranks_per_gpu = 4
walklength = 3
initial_walk = 10
seed = 17
restart_interval = 0
traj_interval = 0
"""
    mpihans.main(dict(NSio.parse_hans_text(base + "iterations = 460\ndirectory = whole\n")))
    os.chdir(tmp_path)
    mpihans.main(dict(NSio.parse_hans_text(base + "iterations = 230\ndirectory = parts\n")))
    os.chdir(tmp_path)
    ns2 = mpihans.main(dict(NSio.parse_hans_text("from_restart = 1\ndirectory = parts\n")))
    os.chdir(tmp_path)
    assert ns2.K == 24 and ns2.nvranks == 4   # ranks_per_gpu came back from the file
    with pytest.raises(ValueError):           # 3 virtual ranks cannot hold a file of 4
        mpihans.main(dict(NSio.parse_hans_text("from_restart = 1\ndirectory = parts\nranks_per_gpu = 3\n")))
    os.chdir(tmp_path)
    whole = open("whole/volumes.txt").read().splitlines()
    parts = open("parts/volumes.txt").read().splitlines()
    assert len(whole) == 461 and whole == parts
    a1, w1 = NSio.open_restart("whole/restart.hdf5")
    a2, w2 = NSio.open_restart("parts/restart.hdf5")
    assert a1["prev_iters"] == a2["prev_iters"] == 460
    for g in w1:
        assert np.array_equal(w1[g][0], w2[g][0]) and np.array_equal(w1[g][1], w2[g][1]) and w1[g][2] == w2[g][2]
    for k in ("dv_max", "dr_max", "dt_max", "dh_max", "dshear", "dstretch"):
        assert a1[k] == a2[k]


def test_initial_config_import(tmp_path, monkeypatch):
    """`initial_config = file.extxyz` (mpihans.py:149-161): every walker starts from the given configuration."""
    from common import make_walkers
    from hans_b200 import NSio, mpihans
    from hans_b200.atoms import Atoms, write_extxyz
    monkeypatch.chdir(tmp_path)
    cfg, H, R, kw = make_walkers("tiny", 1, seed=2)
    write_extxyz(str(tmp_path / "start.extxyz"), Atoms("C12", positions=R[0], pbc=True, cell=H[0]), append=False)
    text = """move_ratio = 1,18,12,0,3,3
nbeads = 2
nchains = 6
nwalkers = 3
ranks_per_gpu = 2
iterations = 1
walklength = 1
initial_walk = 0
directory = imp
seed = 5
initial_config = start.extxyz
"""
    ns = mpihans.main(dict(NSio.parse_hans_text(text)))
    os.chdir(tmp_path)
    attrs, walkers = NSio.open_restart("imp/restart.hdf5")
    # one iteration walked one walker per virtual rank; the other four still hold the imported configuration
    same = [g for g, w in walkers.items() if np.allclose(w[0], R[0], atol=1e-7) and np.allclose(w[1], H[0], atol=1e-7)]
    assert len(walkers) == 6 and len(same) >= 4
    bad = dict(NSio.parse_hans_text(text.replace("nchains = 6", "nchains = 5").replace("directory = imp", "directory = imp2")))
    with pytest.raises(AssertionError):
        mpihans.main(bad)


def test_driver_meets_the_reference_file_contract(tmp_path, monkeypatch):
    """The same input the unmodified reference driver was run on in the build container
    (tests/test_reference_driver_live_cpu.py -> tests/golden/ref_driver_contract.json), through hans_b200.mpihans on the
    B200: same files, same volumes.txt header and line format, same restart attributes (names and types), group names
    and datasets, same trajectory cadence."""
    import json
    import re
    from hans_b200 import NSio, mpihans
    from hans_b200.atoms import read_extxyz
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_driver_contract.json")))
    monkeypatch.chdir(tmp_path)
    mpihans.main(dict(NSio.parse_hans_text(g["input"] + "seed = 3\n")))
    os.chdir(tmp_path)
    assert set(g["files"]) <= set(os.listdir("run1"))
    text = open("run1/volumes.txt").read()
    lines = text.split("\n")
    assert lines[0] == g["volumes_header"] and (lines[-1] == "") == g["trailing_newline"]
    body = [l for l in lines[1:] if l]
    assert len(body) == g["n_volume_lines"] and all(re.match(g["volumes_line_regex"], l) for l in body)
    assert [int(l.split()[0]) for l in body] == list(range(g["n_volume_lines"]))
    frames = read_extxyz("run1/traj.extxyz", ":")
    assert len(frames) == g["traj_frames"]
    attrs, walkers = NSio.open_restart("run1/restart.hdf5")
    assert sorted(walkers) == g["restart_groups"]

    def kind(v):
        a = np.asarray(v)
        return ("str" if a.dtype.kind in "US" else "int" if a.dtype.kind in "iu" else "float" if a.dtype.kind == "f" else a.dtype.kind) + \
               ("" if a.shape == () else str(list(a.shape)))
    for k, t in g["restart_attrs"].items():
        assert k in attrs and kind(attrs[k]) == t, (k, t, attrs.get(k))
    c, u, ctr = walkers["walker_0_0001"]
    assert list(c.shape) == g["restart_datasets"]["coordinates"] and list(u.shape) == g["restart_datasets"]["unitcell"]
