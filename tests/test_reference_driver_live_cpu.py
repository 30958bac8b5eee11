"""The host driver's file contract pinned to the reference EXECUTED HERE: the unmodified /root/reference/mpihans.py and
NesSa/NSio.py run on one fake MPI rank (tests/ref_live_driver.py: hs_alkane on the CPU oracle, ase / mpi4py / h5py stand-ins;
the h5py stand-in stores through hans_b200.h5min), and what they write is compared with the host-side functions of
hans_b200.mpihans / hans_b200.NSio that produce the same files (no GPU needed for those).

Covers SURVEY.md 8(f)1: stdin parser, run directory naming, volumes.txt header and line format, dof, trajectory cadence and
wrapping, restart schema (attributes, group names, datasets) and rotation, restart_cleanup, and the reference's own RESUME path
(mpihans.py:77-89,166-181) reading a restart file written by this package's writer."""
import os
import re

import numpy as np
import pytest

import ref_live

pytestmark = pytest.mark.skipif(not ref_live.reference_available(), reason="/root/reference is not present")

# (no move_ratio / from_restart keys: with them the reference cannot start a fresh run -- `from_restart = 0` is the string
# "0", which MCNS.py:612 takes for true; the default ratio for 4 dimers is 1,12,8,0,3,3)
INPUT = """# comment line
nbeads = 2
nchains = 4
nwalkers = 5
iterations = 230
walklength = 2
initial_walk = 3
time = 1e9
directory = {d}
min_angle = 40.0
"""


@pytest.fixture(scope="module")
def ref_run(tmp_path_factory):
    import ref_live_driver
    tmp = tmp_path_factory.mktemp("refdrv")
    drv = ref_live_driver.LiveDriver(seed=11)
    params, out = drv.run(INPUT.format(d="run1"), str(tmp))
    r = dict(drv=drv, tmp=tmp, params=params, out=out, run=tmp / "run1")
    r["contract"] = _contract(r)   # (before the other tests resume / overwrite the run)
    return r


def test_input_parser_matches(ref_run):
    """NesSa.NSio.read_hans_file (NSio.py:19-67) and hans_b200.NSio.parse_hans_text give the same dictionary."""
    from hans_b200 import NSio
    drv = ref_run["drv"]
    for text in (INPUT.format(d="x"), "nchains = 16\nnbeads=6\nnwalkers = 20\niterations = 800000\nwalklength = 40\ntime = 190000\n"
                 "move_ratio = 3,48,32,0,3,3\nanalyse = 1\nprofile = 0\nbondangle = 111.5\n#x = 3\n\ndv_max = 1.5\n"):
        want = drv.parse(text)
        got = NSio.parse_hans_text(text)
        assert got == want, (got, want)


def test_volumes_file_and_directory(ref_run):
    from hans_b200 import mpihans
    lines = open(ref_run["run"] / "volumes.txt").read().split("\n")
    P = ref_run["params"]
    mr = [1, 12, 8, 0, 3, 3]
    assert lines[0] + "\n" == mpihans.volumes_header(P["nwalkers"] * 1, mpihans.compute_dof(P, mr), P["nchains"])   # mpihans.py:202
    body = [l for l in lines[1:] if l]
    assert len(body) == 230
    vols = []
    for i, l in enumerate(body):
        it, v1, v2 = l.split()
        assert int(it) == i and v1 == v2
        assert l + "\n" == mpihans.volumes_line(i, float(v1))                                                        # mpihans.py:221
        vols.append(float(v1))
    assert (np.diff(vols) <= 1e-9).all()      # the ceiling never rises
    # gen_prefix directory naming (mpihans.py:52-58)
    drv, tmp = ref_run["drv"], ref_run["tmp"]
    cwd = os.getcwd()
    os.chdir(tmp)
    try:
        want = mpihans.gen_directory(drv.parse(INPUT.format(d="gen_prefix")), 1)
    finally:
        os.chdir(cwd)
    drv.run(INPUT.format(d="gen_prefix").replace("iterations = 230", "iterations = 3"), str(tmp))
    assert os.path.isdir(tmp / want) and want == "NeSa_4_2mer.5.2.1/"
    # the stdout log: one line per step adaptation, `i vol_max [6 rates]` (mpihans.py:252-253), every K // 2 iterations
    adj = [l for l in ref_run["out"].splitlines() if re.match(r"^\d+ \d+\.\d+ \[", l)]
    assert [int(l.split()[0]) for l in adj] == list(range(0, 230, max(5 // 2, 1)))


def test_trajectory_cadence_and_wrapping(ref_run):
    """traj.extxyz: the OLD max walker every 100 iterations (mpihans.py:38,232-234), wrapped (NSio.py:121-124)."""
    drv = ref_run["drv"]
    writes = [c for c in drv.aseio.calls if c[0] == "write" and c[1] == "traj.extxyz"]
    key = os.path.abspath(ref_run["run"] / "traj.extxyz")
    frames = drv.aseio.files[key]
    assert len(frames) == 3 and all(c[2] is True and c[3] == 1 for c in writes[:3])     # iterations 0, 100, 200, appended
    vols = [float(l.split()[1]) for l in open(ref_run["run"] / "volumes.txt").read().split("\n")[1:] if l]
    for k, (cell, pos, wrapped) in enumerate(frames):
        assert wrapped and abs(abs(np.linalg.det(cell)) - vols[100 * k]) < 1e-9 * vols[100 * k]
        frac = np.linalg.solve(cell.T, pos.T).T
        assert (frac >= -1e-7 - 1e-12).all() and (frac < 1 - 1e-7 + 1e-12).all()
    # this package's writer makes the same frame from the same configuration
    from hans_b200 import NSio
    from hans_b200.atoms import read_extxyz
    cell, pos, _ = frames[1]
    fn = str(ref_run["tmp"] / "mine.extxyz")
    NSio.write_to_extxyz(cell, pos + cell[0] * 2 - cell[2], filename=fn)   # moved by lattice vectors: wraps back
    at = read_extxyz(fn)
    assert np.allclose(at.positions, pos, atol=2e-8) and np.allclose(at.cell, cell, atol=0, rtol=0)


def test_restart_schema_matches_and_reference_resumes_from_our_file(ref_run):
    """restart.hdf5 as the reference creates it (NSio.py:69-111) against hans_b200.NSio.write_to_restart on the same state:
    attributes, group names, datasets.  Then the reference's own resume path (from_restart = 1) runs on a restart file
    WRITTEN BY THIS PACKAGE and continues volumes.txt."""
    import ref_live_driver
    from hans_b200 import NSio, h5min
    drv, run = ref_run["drv"], ref_run["run"]
    assert ("w", "restart.hdf5") in ref_live_driver.FakeH5File.log
    attrs, groups = h5min.read(str(run / "restart.hdf5"))                  # (the reference's calls, stored through h5min)
    assert sorted(groups) == [f"walker_0_{i:04d}" for i in range(1, 6)]                                   # NSio.py:87
    assert all(sorted(g) == ["coordinates", "unitcell"] and g["coordinates"].shape == (8, 3) and g["unitcell"].shape == (3, 3)
               for g in groups.values())
    assert int(attrs["prev_iters"]) == 230 and list(attrs["move_ratio"]) == [1, 12, 8, 0, 3, 3]
    for k in ("dv_max", "dr_max", "dt_max", "dh_max", "dshear", "dstretch", "nchains", "nbeads", "nwalkers", "walklength",
              "iterations", "min_angle", "min_aspect_ratio", "bondlength", "bondangle", "initial_walk", "restart_file", "time"):
        assert k in attrs, k
    # our writer on the same walkers and parameters
    H = np.stack([groups[f"walker_0_{i:04d}"]["unitcell"] for i in range(1, 6)])
    R = np.stack([groups[f"walker_0_{i:04d}"]["coordinates"] for i in range(1, 6)])
    P = dict(ref_run["params"])
    steps = {k: float(attrs[k]) for k in ("dv_max", "dr_max", "dt_max", "dh_max", "dshear", "dstretch")}
    mine = str(ref_run["tmp"] / "mine.hdf5")
    NSio.write_to_restart(P, [(H, R, np.arange(5) * 7)], filename=mine, i=229, steps=steps, nwalkers=5)
    a2, g2 = h5min.read(mine)
    assert sorted(g2) == sorted(groups)
    for g in groups:
        assert np.array_equal(g2[g]["coordinates"], groups[g]["coordinates"]) and np.array_equal(g2[g]["unitcell"], groups[g]["unitcell"])
        assert sorted(g2[g]) == ["coordinates", "philox_ctr", "unitcell"]                  # (+ our own extra dataset)
    for k in attrs:
        v1, v2 = attrs[k], a2[k]
        assert np.array_equal(np.asarray(v1), np.asarray(v2)) or str(v1) == str(v2), (k, v1, v2)
    # the reference resumes from OUR file: swap it in, rerun with from_restart = 1 (mpihans.py:43-94,166-181)
    os.replace(mine, run / "restart.hdf5")
    n_before = len(ref_live_driver.FakeH5File.log)
    params, out = drv.run("from_restart = 1\ndirectory = run1\ntime = 1e9\n", str(ref_run["tmp"]))
    assert ("r", "restart.hdf5") in ref_live_driver.FakeH5File.log[n_before:]
    assert "Attempting to restart from run1" in out
    lines = [l for l in open(run / "volumes.txt").read().split("\n") if l]
    assert len(lines) == 1 + 460 and int(lines[-1].split()[0]) == 459                  # iterations restored from the file: 230 more
    vols = np.array([float(l.split()[1]) for l in lines[1:]])
    assert (np.diff(vols) <= 1e-9).all()
    assert os.path.exists(run / "restart_backup.hdf5")                                 # rotation, mpihans.py:294-297


def test_restart_cleanup_matches(ref_run, tmp_path):
    """NSio.restart_cleanup (NSio.py:127-144) of the reference and of this package truncate volumes.txt alike."""
    from hans_b200 import NSio
    drv = ref_run["drv"]
    text = "40 1 64 False 16 \n" + "".join(f"{i} {1000 - i:.13f} {1000 - i:.13f} \n" for i in range(350))
    outs = []
    for who in ("ref", "mine"):
        d = tmp_path / who
        d.mkdir()
        (d / "volumes.txt").write_text(text)
        cwd = os.getcwd()
        os.chdir(d)
        try:
            if who == "ref":
                key = os.path.abspath("traj.extxyz")
                drv.aseio.files[key] = [(np.eye(3) * (10 - k), np.zeros((2, 3)), True) for k in range(4)]
                drv.NSio.restart_cleanup(dict(prev_iters=250), 100)
                nframes = len(drv.aseio.files[key])
            else:
                for k in range(4):
                    NSio.write_to_extxyz(np.eye(3) * (10 - k), np.zeros((2, 3)), "traj.extxyz")
                NSio.restart_cleanup(dict(prev_iters=250), traj_interval=100)
                from hans_b200.atoms import read_extxyz
                nframes = len(read_extxyz("traj.extxyz", ":"))
        finally:
            os.chdir(cwd)
        outs.append(((d / "volumes.txt").read_text(), nframes))
    assert outs[0] == outs[1] and outs[0][1] == 2 and len(outs[0][0].splitlines()) == 251


def _contract(ref_run):
    """What the reference's run left behind, reduced to the file contract (values that depend on random draws left out)."""
    from hans_b200 import h5min
    run = ref_run["run"]
    lines = open(run / "volumes.txt").read().split("\n")
    attrs, groups = h5min.read(str(run / "restart.hdf5"))

    def kind(v):
        a = np.asarray(v)
        return ("str" if a.dtype.kind in "US" else "int" if a.dtype.kind in "iu" else "float" if a.dtype.kind == "f" else a.dtype.kind) + \
               ("" if a.shape == () else str(list(a.shape)))
    return dict(input=INPUT.format(d="run1"), volumes_header=lines[0], n_volume_lines=len([l for l in lines[1:] if l]),
                volumes_line_regex=r"^\d+ \d+\.\d{13} \d+\.\d{13} $", trailing_newline=lines[-1] == "",
                restart_attrs={k: kind(v) for k, v in sorted(attrs.items())}, restart_groups=sorted(groups),
                restart_datasets={n: list(a.shape) for n, a in sorted(groups["walker_0_0001"].items())},
                traj_frames=len(ref_run["drv"].aseio.files[os.path.abspath(run / "traj.extxyz")]), traj_iterations=[0, 100, 200],
                files=sorted(os.listdir(run)) + ["traj.extxyz"], adjust_log_iterations=list(range(0, 230, 2)))


def test_driver_contract_golden_is_current(ref_run):
    """tests/golden/ref_driver_contract.json (the file contract as the reference's own run exhibits it; used by
    tests/test_gpu_driver.py on the B200, where the reference is absent) is what this run produces."""
    import json
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_driver_contract.json")
    got = ref_run["contract"]
    for l in open(ref_run["run"] / "volumes.txt").read().split("\n")[1:]:
        assert l == "" or re.match(got["volumes_line_regex"], l), l
    if os.environ.get("HANS_WRITE_GOLDEN") == "1":
        json.dump(got, open(path, "w"), indent=1)
    want = json.load(open(path))
    assert got == want
