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
    c, u = walkers["walker_1_0002"]
    assert c.shape == (96, 3) and u.shape == (3, 3) and np.array_equal(c, states[1][1][1]) and np.array_equal(u, states[1][0][1])
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
