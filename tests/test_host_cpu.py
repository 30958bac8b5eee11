"""CPU tests of the host layer: input parsing, file formats, the NS bookkeeping rules."""
import os

import numpy as np
import pytest

README_INPUT = """move_ratio = 3,48,32,0,3,3
nbeads = 6
nchains = 16
nwalkers = 20
iterations = 800000
walklength = 40
time = 190000
analyse = 1
"""
# the reference's committed input.txt (no `time`, no `from_restart`, mixed spacing)
INPUT_TXT = """nwalkers = 125
nchains = 32
nbeads=2
bondlength = 0.4
bondangle=109.7
iterations=100
walklength = 320
min_angle = 60
min_aspect_ratio=0.8
"""


def test_parse_readme_and_input_txt():
    from hans_b200 import NSio
    p = NSio.parse_hans_text(README_INPUT)
    assert p["nbeads"] == 6 and p["nchains"] == 16 and p["nwalkers"] == 20 and p["walklength"] == 40
    assert p["iterations"] == "800000" and p["move_ratio"] == "3,48,32,0,3,3"      # stay strings, NSio.py:35-46
    assert p["time"] == 190000.0 and p["analyse"] == 1
    # defaults NSio.py:47-62
    assert p["initial_walk"] == 1000 and p["bondlength"] == 0.4 and p["bondangle"] == 109.7
    assert p["min_angle"] == 45.0 and p["min_aspect_ratio"] == 0.8 and p["restart_file"] == "restart.hdf5"
    q = NSio.parse_hans_text(INPUT_TXT + "# a comment\n\nprofile = 1\n")
    assert q["nwalkers"] == 125 and q["nbeads"] == 2 and q["min_angle"] == 60.0 and q["profile"] is True
    assert "time" not in q and "from_restart" not in q
    with pytest.raises(SystemExit):
        NSio.parse_hans_text("")


def test_default_move_ratio_table():
    from hans_b200 import MCNS as NS
    # SURVEY.md section 8 size table / MCNS.py:619-626
    assert list(NS.default_move_ratio(dict(nchains=64, nbeads=1))) == [1, 192, 0, 0, 3, 3]
    assert list(NS.default_move_ratio(dict(nchains=128, nbeads=6))) == [1, 384, 256, 384, 3, 3]
    assert list(NS.default_move_ratio(dict(nchains=32, nbeads=2))) == [1, 96, 64, 0, 3, 3]
    assert list(NS.default_move_ratio(dict(nchains=16, nbeads=6, move_ratio="3,48,32,0,3,3"))) == [3, 48, 32, 0, 3, 3]
    with pytest.raises(Exception):
        NS.default_move_ratio(dict(nchains=1, nbeads=1, move_ratio="1,2,3"))


def test_volumes_format_dof_and_directory(tmp_path, monkeypatch):
    from hans_b200 import mpihans
    assert mpihans.volumes_header(160, 64, 16) == "160 1 64 False 16 \n"            # mpihans.py:202
    assert mpihans.volumes_line(7, 960.2038395943907) == "7 960.2038395943907 960.2038395943907 \n"  # :221
    P = dict(nchains=16, nbeads=6)
    assert mpihans.compute_dof(P, [3, 48, 32, 0, 3, 3]) == 16 + 48                  # mpihans.py:189-196
    assert mpihans.compute_dof(dict(nchains=32, nbeads=2), [1, 96, 64, 0, 3, 3]) == 32 + 64
    assert mpihans.compute_dof(P, [0, 0, 1, 0, 0, 0]) == 0
    monkeypatch.chdir(tmp_path)
    S = dict(nchains=16, nbeads=6, nwalkers=20, walklength=40)
    assert mpihans.gen_directory(S, 4) == "NeSa_16_6mer.80.160.1/"                  # mpihans.py:54-58
    os.mkdir("NeSa_16_6mer.80.160.1")
    assert mpihans.gen_directory(S, 4) == "NeSa_16_6mer.80.160.2/"


def test_chunk_boundaries():
    from hans_b200.mpihans import chunk_end
    # after iteration 0 the reference adjusts (0 % interval == 0) and writes a trajectory frame
    assert chunk_end(0, 999, 80, 100, 50000, 4096) == 0
    assert chunk_end(1, 999, 80, 100, 50000, 4096) == 80
    assert chunk_end(81, 999, 80, 100, 50000, 4096) == 100
    assert chunk_end(161, 165, 80, 100, 50000, 4096) == 165                        # end of run
    assert chunk_end(49901, 10 ** 6, 10 ** 5, 0, 50000, 10 ** 6) == 49999          # restart after (i+1) % 50000 == 0
    assert chunk_end(1, 10 ** 6, 10 ** 5, 0, 10 ** 6, 64) == 64                    # log ring bound


def test_update_steps_rule():
    from hans_b200.ns import STEP_MIN, update_steps
    s = dict(dv_max=30.0, dr_max=2.0, dt_max=0.43, dh_max=0.4, dshear=1.0, dstretch=2e-5)
    mr = [1, 1, 1, 0, 1, 1]
    out = update_steps(s, [0.6, 0.1, 0.3, 0.9, 0.51, 0.19], mr)                    # MCNS.py:687-716
    assert out["dv_max"] == 50.0            # doubled, capped at 50 (mpihans.py:24)
    assert out["dr_max"] == 1.0             # halved
    assert out["dt_max"] == 0.43            # inside [0.2, 0.5]: unchanged
    assert out["dh_max"] == 0.4             # move type disabled: untouched
    assert out["dshear"] == 2.0             # doubled, no cap
    assert out["dstretch"] == 1e-5          # floored at 1e-5 (mpihans.py:31-32)
    assert update_steps(dict(s, dr_max=1.5e-10), [0.3, 0.0, 0.3, 0, 0.3, 0.3], mr)["dr_max"] == STEP_MIN[1]


def test_resolve_host_tie_rule_and_clone_source():
    from hans_b200 import philox
    from hans_b200.ns import resolve_host
    rec = np.zeros((4, 15))
    rec[:, 0] = [5.0, 7.0, 7.0, 6.0]
    rec[:, 1] = [3, 9, 1, 2]
    vmax, r, i, sr, si = resolve_host(rec, seed=42, iteration=17, nlocal=10)
    assert (vmax, r, i) == (7.0, 1, 9)                                             # lowest rank wins a tie
    assert (sr, si) == philox.clone_source(42, 17, 10, 4) and 0 <= sr < 4 and 0 <= si < 10
    # uniform over all K walkers including the max one (mpihans.py:220)
    g = np.array([philox.ns_draw(1, it, 0, 0) % 40 for it in range(4000)])
    assert np.bincount(g, minlength=40).min() > 50


def test_extxyz_roundtrip_and_wrap(tmp_path):
    from hans_b200 import NSio
    from hans_b200.atoms import Atoms, read_extxyz
    H = np.array([[4.0, 0.0, 0.0], [1.0, 3.0, 0.0], [0.5, 0.2, 5.0]])
    R = np.array([[9.0, -1.0, 2.0], [0.1, 0.2, 0.3], [-4.0, 7.0, 11.0]])
    fn = str(tmp_path / "traj.extxyz")
    NSio.write_to_extxyz(H, R, fn)
    NSio.write_to_extxyz(H * 2, R, fn)
    text = open(fn).read().splitlines()
    assert text[0] == "3"
    assert text[1] == 'Lattice="4.0 0.0 0.0 1.0 3.0 0.0 0.5 0.2 5.0" Properties=species:S:1:pos:R:3 pbc="T T T"'
    assert text[3].startswith("C ") and len(text[3].split()) == 4
    frames = read_extxyz(fn, ":")
    assert len(frames) == 2 and np.allclose(frames[0].cell, H) and np.allclose(frames[1].cell, 2 * H)
    frac = np.linalg.solve(H.T, frames[0].positions.T).T
    assert (frac >= -1e-7 - 1e-12).all() and (frac < 1 - 1e-7 + 1e-12).all()          # ase wrap convention
    d = np.linalg.solve(H.T, (frames[0].positions - R).T).T
    assert np.allclose(d, np.round(d), atol=1e-7)                                    # moved by lattice vectors only
    assert read_extxyz(fn).get_volume() == pytest.approx(8 * 60.0)
    a = Atoms("C2", positions=[[0, 0, 0], [1, 1, 1]], cell=[2.0, 3.0, 4.0])
    assert a.cell.shape == (3, 3) and a.cell[1, 1] == 3.0


def test_restart_container_and_cleanup(tmp_path, monkeypatch):
    from hans_b200 import NSio
    monkeypatch.chdir(tmp_path)
    P = NSio.parse_hans_text(README_INPUT)
    P["move_ratio"] = np.array([3.0, 48, 32, 0, 3, 3])
    rng = np.random.default_rng(0)
    states = [(rng.normal(size=(2, 3, 3)), rng.normal(size=(2, 96, 3))) for _ in range(2)]
    steps = dict(dv_max=1.0, dr_max=0.5, dt_max=0.2, dh_max=0.1, dshear=0.3, dstretch=0.4)
    NSio.write_to_restart(P, states, filename="restart.hdf5", i=49999, steps=steps)
    assert NSio.restart_exists("restart.hdf5")
    attrs, walkers = NSio.open_restart("restart.hdf5")
    assert attrs["prev_iters"] == 50000 and attrs["dshear"] == 0.3 and attrs["nchains"] == 16          # NSio.py:74-83
    assert attrs["iterations"] == "800000" and list(attrs["move_ratio"]) == [3, 48, 32, 0, 3, 3]
    assert sorted(walkers) == ["walker_0_0001", "walker_0_0002", "walker_1_0001", "walker_1_0002"]    # NSio.py:87
    c, u, ctr = walkers["walker_1_0002"]
    assert ctr is None   # no counters were passed
    assert c.shape == (96, 3) and u.shape == (3, 3) and np.array_equal(c, states[1][1][1]) and np.array_equal(u, states[1][0][1])
    # virtual ranks: one GPU holding 4 walkers = 2 ranks x nwalkers 2 writes the SAME groups as two GPUs with 2 each
    one_gpu = [(np.concatenate([states[0][0], states[1][0]]), np.concatenate([states[0][1], states[1][1]]), np.arange(4) * 1000)]
    NSio.write_to_restart(P, one_gpu, filename="restart_v.hdf5", i=49999, steps=steps, nwalkers=2)
    _, wv = NSio.open_restart("restart_v.hdf5")
    assert sorted(wv) == sorted(walkers)
    for g in walkers:
        assert np.array_equal(wv[g][0], walkers[g][0]) and np.array_equal(wv[g][1], walkers[g][1])
    assert [wv[g][2] for g in sorted(wv)] == [0.0, 1000.0, 2000.0, 3000.0]   # Philox counters (dataset philox_ctr)
    # restart_cleanup, NSio.py:127-144
    with open("volumes.txt", "w") as f:
        f.write("40 1 64 False 16 \n")
        for i in range(350):
            f.write(f"{i} {1000 - i:.13f} {1000 - i:.13f} \n")
    for k in range(4):
        NSio.write_to_extxyz(np.eye(3) * (10 - k), np.zeros((2, 3)), "traj.extxyz")
    NSio.restart_cleanup(dict(prev_iters=250), traj_interval=100)
    assert len(open("volumes.txt").read().splitlines()) == 251
    from hans_b200.atoms import read_extxyz
    assert len(read_extxyz("traj.extxyz", ":")) == 2


def test_h5min_round_trip_and_structure(tmp_path):
    """restart.hdf5 without h5py: the dependency-free HDF5 writer (hans_b200/h5min.py).  Round trip through its
    own reader, and the structural rules libhdf5's reader enforces: end-of-file address = file size, nodes at
    full size, names sorted, group B-tree keys bounding the names of child i as (key[i], key[i+1]], siblings
    linked, heap free list = 1 (H5HL_FREE_NULL)."""
    import struct
    from hans_b200 import h5min
    rng = np.random.default_rng(3)
    ngroups = 600  # 75 symbol table nodes -> 3 leaf-level B-tree nodes -> a second B-tree level
    groups = {f"walker_{i % 3}_{i // 3 + 1:04d}": {"coordinates": rng.normal(size=(12, 3)), "unitcell": rng.normal(size=(3, 3))}
              for i in range(ngroups)}
    attrs = {"nchains": 16, "bondlength": 0.4, "iterations": "800000", "directory": "NeSa_16_6mer.20.40.1",
             "move_ratio": np.array([3.0, 48, 32, 0, 3, 3]), "profile": False, "prev_iters": 50000, "time": float("inf"),
             "seed": 2 ** 40 + 7, "empty": "", "unicode": "σ = 1"}
    fn = str(tmp_path / "restart.hdf5")
    h5min.write(fn, attrs, groups)
    assert h5min.is_hdf5(fn)
    a2, g2 = h5min.read(fn)
    assert set(a2) == set(attrs)
    assert a2["nchains"] == 16 and a2["bondlength"] == 0.4 and a2["iterations"] == "800000" and a2["seed"] == 2 ** 40 + 7
    assert a2["profile"] == 0 and a2["time"] == float("inf") and a2["empty"] == "" and a2["unicode"] == "σ = 1"
    assert list(a2["move_ratio"]) == [3, 48, 32, 0, 3, 3] and a2["directory"] == attrs["directory"]
    assert sorted(g2) == sorted(groups)
    for g in groups:
        for d in ("coordinates", "unitcell"):
            assert g2[g][d].dtype == np.float64 and np.array_equal(g2[g][d], groups[g][d])
    # ---- structure ----
    raw = open(fn, "rb").read()
    assert raw[:8] == b"\x89HDF\r\n\x1a\n" and raw[8] == 0 and raw[13] == 8 and raw[14] == 8
    leaf_k, int_k = struct.unpack_from("<HH", raw, 16)
    assert (leaf_k, int_k) == (4, 16)
    base, free, eof, drv = struct.unpack_from("<QQQQ", raw, 24)
    assert base == 0 and eof == len(raw) and free == drv == 0xFFFFFFFFFFFFFFFF
    name_off, root_hdr, cache, _ = struct.unpack_from("<QQII", raw, 56)
    bt_c, heap_c = struct.unpack_from("<QQ", raw, 80)
    r = h5min._Reader(raw)
    stab = [m for t, m in r.messages(root_hdr) if t == 0x0011][0]
    assert cache == 1 and struct.unpack("<QQ", stab[:16]) == (bt_c, heap_c) and name_off == 0
    assert raw[heap_c:heap_c + 4] == b"HEAP"
    seg_size, free_head, seg = struct.unpack_from("<QQQ", raw, heap_c + 8)
    assert free_head == 1 and seg % 8 == 0 and seg_size % 8 == 0 and raw[seg] == 0

    def name(off):
        return r.heap_name(heap_c, off).encode()

    levels = {}

    def walk(addr, lo, hi):
        """every name below `addr` must lie in (lo, hi]"""
        assert addr % 8 == 0
        if raw[addr:addr + 4] == b"SNOD":
            assert addr + 8 + 40 * 8 <= len(raw)
            n = struct.unpack_from("<H", raw, addr + 6)[0]
            names = [name(struct.unpack_from("<Q", raw, addr + 8 + 40 * i)[0]) for i in range(n)]
            assert 1 <= n <= 8 and names == sorted(names) and lo < names[0] and names[-1] <= hi
            for i in range(n):  # sub-groups carry their B-tree / heap in the scratch pad
                _, hdr, ctype, _ = struct.unpack_from("<QQII", raw, addr + 8 + 40 * i)
                assert ctype == 1
                sub = [m for t, m in r.messages(hdr) if t == 0x0011][0]
                assert raw[addr + 8 + 40 * i + 24:addr + 8 + 40 * i + 40] == sub[:16]
            return n
        assert raw[addr:addr + 4] == b"TREE" and addr + 24 + 8 * 32 + 8 * 33 <= len(raw)
        ntype, lvl, used, left, right = struct.unpack_from("<BBHQQ", raw, addr + 4)
        assert ntype == 0 and 1 <= used <= 32
        levels.setdefault(lvl, []).append((addr, left, right))
        keys = [name(struct.unpack_from("<Q", raw, addr + 24 + 16 * i)[0]) for i in range(used + 1)]
        assert keys[0] == lo and keys[-1] == hi and keys == sorted(keys)
        total = 0
        for i in range(used):
            child = struct.unpack_from("<Q", raw, addr + 24 + 8 + 16 * i)[0]
            total += walk(child, keys[i], keys[i + 1])
        return total

    root_used = struct.unpack_from("<H", raw, bt_c + 6)[0]
    top_key = name(struct.unpack_from("<Q", raw, bt_c + 24 + 16 * root_used)[0])
    assert walk(bt_c, b"", top_key) == ngroups and top_key == max(g.encode() for g in groups)
    assert sorted(levels) == [0, 1] and len(levels[0]) == 3 and len(levels[1]) == 1
    undef = 0xFFFFFFFFFFFFFFFF
    for lvl, nodes in levels.items():  # siblings, left to right
        assert nodes[0][1] == undef and nodes[-1][2] == undef
        for (a0, _, r0), (a1, l1, _) in zip(nodes, nodes[1:]):
            assert r0 == a1 and l1 == a0
    # a dataset header: dataspace, datatype (IEEE float64 little endian), fill value, contiguous layout
    hdr = r.links(r.links(root_hdr)["walker_0_0001"])["coordinates"]
    ms = dict(r.messages(hdr))
    assert ms[0x0003][:20] == bytes.fromhex("11203f00080000000000400034" "0b0034ff030000")
    assert ms[0x0001][:24] == struct.pack("<BBB5xQQ", 1, 2, 0, 12, 3) and ms[0x0008][:2] == b"\x03\x01"
    addr, size = struct.unpack_from("<QQ", ms[0x0008], 2)
    assert size == 12 * 3 * 8 and addr % 8 == 0 and addr + size <= len(raw)
    # an empty group and an empty file are still well formed
    h5min.write(fn, {}, {"lonely": {}})
    a3, g3 = h5min.read(fn)
    assert a3 == {} and g3 == {"lonely": {}}


def test_h5min_file_opens_with_h5py(tmp_path):
    """Cross-check against libhdf5 where h5py exists (it is absent from the build image, so this is skipped there)."""
    h5py = pytest.importorskip("h5py")
    from hans_b200 import h5min
    rng = np.random.default_rng(4)
    groups = {f"walker_0_{i + 1:04d}": {"coordinates": rng.normal(size=(6, 3)), "unitcell": rng.normal(size=(3, 3))} for i in range(300)}
    fn = str(tmp_path / "restart.hdf5")
    h5min.write(fn, {"nchains": 16, "bondlength": 0.4, "iterations": "800000", "move_ratio": np.arange(6.0)}, groups)
    with h5py.File(fn, "r") as f:
        assert sorted(f) == sorted(groups) and f.attrs["nchains"] == 16 and f.attrs["bondlength"] == 0.4
        assert list(f.attrs["move_ratio"]) == list(range(6))
        assert np.array_equal(f["walker_0_0007"]["unitcell"][:], groups["walker_0_0007"]["unitcell"])


def test_h5min_reader_on_a_libhdf5_written_file():
    """The one libhdf5-written file in the image: scipy's MATLAB v7.3 test file (superblock 0 behind a 512-byte user
    block, old-style root group, one contiguous float64 dataset with a string attribute).  The reader shares every
    structure with the writer (superblock, local heap, B-tree keys, symbol table node, version 1 object header,
    dataspace / datatype / attribute messages), so parsing it pins the writer's understanding of the format."""
    import scipy.io
    fn = os.path.join(os.path.dirname(scipy.io.__file__), "matlab", "tests", "data", "testhdf5_7.4_GLNX86.mat")
    if not os.path.exists(fn):
        pytest.skip("scipy's test data is not installed")
    import struct
    from hans_b200 import h5min
    raw = open(fn, "rb").read()
    r = h5min._Reader(raw)
    links = r.links(r.root_header)
    assert list(links) == ["testdouble"]
    a = r.dataset(links["testdouble"])
    assert a.shape == (9, 1) and np.allclose(a[:, 0], np.arange(9) * np.pi / 4)
    assert r.attributes(links["testdouble"]) == {"MATLAB_class": "double"}
    # what libhdf5 wrote where the writer has to make the same choices
    d = r.d
    bt, heap = struct.unpack_from("<QQ", d, 80)
    used = struct.unpack_from("<H", d, bt + 6)[0]
    k0, child, k1 = struct.unpack_from("<QQQ", d, bt + 24)
    assert used == 1 and r.heap_name(heap, k0) == "" and r.heap_name(heap, k1) == "testdouble"     # (key[i], key[i+1]]
    free_head = struct.unpack_from("<Q", d, heap + 16)[0]
    seg = struct.unpack_from("<Q", d, heap + 24)[0]
    assert struct.unpack_from("<Q", d, seg + free_head)[0] == 1                                     # H5HL_FREE_NULL ends the free list
    ms = dict(r.messages(links["testdouble"]))
    assert ms[0x0003][:20] == h5min._dt_f64() and ms[0x0001][:24] == h5min._dataspace((9, 1))


def test_window_settle_rule_equals_serial_order():
    """The rule the warp-per-move kernel settles a window with (hans_b200/csrc/hans_team.cuh), on a toy system where
    it can be enumerated: rods on a ring, a move shifts one rod, a move is accepted iff the shifted rod overlaps no
    other rod.  Window = consecutive moves of DISTINCT rods; every move is evaluated against the state at the START of
    the window, split into `base` (rods no earlier move of the window touches), `old[j]` / `new[j]` (the rod of the
    earlier move j at its old / trial position); then in order: overlap_k = base_k or any_j<k (committed_j ? new_kj :
    old_kj).  Must equal one-move-at-a-time evaluation, for every window length."""
    rng = np.random.default_rng(11)
    L, n = 40.0, 12

    def ov(a, b):
        d = abs(a - b) % L
        return min(d, L - d) < 1.0

    for W in (2, 4, 8):
        x = np.arange(n) * (L / n) + rng.uniform(-0.5, 0.5, n)
        moves = [(int(rng.integers(n)), float(rng.normal(scale=2.5))) for _ in range(4000)]
        xs = x.copy()  # serial reference
        serial = []
        for c, dx in moves:
            t = (xs[c] + dx) % L
            ok = not any(ov(t, xs[j]) for j in range(n) if j != c)
            serial.append(ok)
            if ok:
                xs[c] = t
        xw, got, q, nwin = x.copy(), [], 0, 0
        while q < len(moves):
            nw = 1
            while nw < W and q + nw < len(moves) and moves[q + nw][0] not in [m[0] for m in moves[q:q + nw]]:
                nw += 1
            win = moves[q:q + nw]
            trial = [(xw[c] + dx) % L for c, dx in win]
            committed = []
            for k, (c, _) in enumerate(win):
                early = [m[0] for m in win[:k]]
                base = any(ov(trial[k], xw[j]) for j in range(n) if j != c and j not in early)
                old = [ov(trial[k], xw[cj]) for cj in early]
                new = [ov(trial[k], trial[j]) for j in range(k)]
                over = base or any(new[j] if committed[j] else old[j] for j in range(k))
                committed.append(not over)
            for k, (c, _) in enumerate(win):
                if committed[k]:
                    xw[c] = trial[k]
            got += committed
            q += nw
            nwin += 1
        assert got == serial and np.array_equal(xw, xs)
        assert 0.2 < np.mean(serial) < 0.95 and nwin < len(moves) / min(W, 3) * 1.6


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm): exit 0, exactly ONE line on stdout, a
    JSON object with the contract's keys and the same workload string the GPU arm reports for that workload."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout[:500]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "mc_move_attempts_per_s" and d["unit"] == "moves/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 3
    assert d["value"] > 0 and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] in ("port", "reference")
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    sys.path.insert(0, root)
    import bench
    assert d["config"]["workload"].startswith("C3") and d["config"]["workload"] == bench.workload_string("C3", bench.WORKLOADS["C3"][3])
