"""The oracle pinned to the reference EXECUTED HERE: /root/reference/NesSa/MCNS.py is imported
unmodified (tests/ref_live.py supplies stand-ins for the absent hs_alkane / ase / mpi4py) and its
own MC_run, box_shear_step, box_stretch_step, min_aspect_ratio, min_angle, adjust_mc_steps,
default_move_ratio run on scripted random draws; results are compared with oracle/hans_oracle.c
(ho_apply_proposal & co) on the same draws.

What is bit-exact and what is not (stated, not hidden):
  * move selection, chain selection, draw order and count, accept / revert / ceiling logic, rate
    bookkeeping, volume / translate / rotate / dihedral outcomes, the shear step's delta_H: bit-exact
    (np.dot of 3-vectors in this image is an FMA chain; the oracle's shear path uses the same order);
  * the stretch step multiplies by np.exp(rv): libm's exp, which the device cannot reproduce bit for
    bit; the oracle / GPU use the deterministic hd_exp (<= 2 ulp from libm, test_oracle_cpu.py).  With
    the reference's `np.exp` swapped for that routine -- the ONLY substitution -- whole runs are
    bit-identical; with numpy's own exp, decisions are identical and states agree to <= 1e-13 relative.

These tests need /root/reference and are skipped where it is absent (the GPU box); the vectors they
produce travel as tests/golden/ref_live_*.{npz,json} (tests/golden/make_ref_live_golden.py).
"""
import math
import sys

import numpy as np
import pytest

import ref_live
from common import SHAPES
from oracle import oracle as ho

pytestmark = pytest.mark.skipif(not ref_live.reference_available(), reason="/root/reference is not present")

_LIVE = {}


def live(det_exp=False):
    if det_exp not in _LIVE:
        ov = {"exp": (lambda x: ho.det_exp(float(x)))} if det_exp else None
        _LIVE[det_exp] = ref_live.LiveReference(np_overrides=ov)
    return _LIVE[det_exp]


def start_state(shape, seed, presweeps=2, **kw):
    nchains, nbeads, mr = SHAPES[shape]
    cfg = ho.make_cfg(nchains, nbeads, mr, **kw)
    H, R = ho.grow_walkers(cfg, 1, seed)
    if presweeps:
        ctrs = np.zeros(1, dtype=np.uint64)
        ho.mc_run_many(cfg, H, R, [0], presweeps, 1e300, seed + 1, [0], ctrs)
    return cfg, mr, H[0].copy(), R[0].copy()


def run_both(L, shape, seed, sweeps, limit_factor=None, **kw):
    cfg, mr, H0, R0 = start_state(shape, seed, **kw)
    mps = ho.moves_per_sweep(cfg.nbeads, cfg.nchains)
    props = ho.gen_proposals(cfg, 1000 + seed, 0, 0, sweeps * mps)
    limit = sys.float_info.max if limit_factor is None else ho.volume(H0) * limit_factor
    L.setup(cfg.nchains, cfg.nbeads, H0, R0)
    vol, rates = L.mc_run(cfg, props, move_ratio=mr, volume_limit=limit, min_ang=45.0, bump_stretch=True)
    assert L.drained(), "the reference drew fewer random numbers than SURVEY.md 10.1 says"
    Ho, Ro = H0.copy(), R0.copy()
    dec = ho.replay(cfg, Ho, Ro, props, volume_limit=limit)
    att = np.bincount(props["type"], minlength=6).astype(float)
    acc = np.bincount(props["type"], weights=dec["accepted"], minlength=6)
    o_rates = acc / np.where(att == 0, 1, att)
    return dict(cfg=cfg, props=props, dec=dec, vol=vol, rates=rates, o_rates=o_rates, Href=L.alk.H[0].copy(),
                Rref=L.alk.R[0].copy(), Ho=Ho, Ro=Ro, limit=limit)


CASES = [("tiny1", 3, 6, None), ("tiny", 4, 5, 1.01), ("C1", 5, 3, 1.005), ("C2", 6, 4, 1.002), ("p7", 7, 3, None),
         ("t3", 8, 4, 1.01), ("h15", 9, 2, 1.003)]


@pytest.mark.parametrize("shape,seed,sweeps,lf", CASES)
def test_reference_mc_run_decisions_equal_oracle(shape, seed, sweeps, lf):
    """Unmodified reference, numpy's own arithmetic throughout."""
    r = run_both(live(False), shape, seed, sweeps, lf, dshear=0.4, dstretch=0.08)
    assert np.array_equal(r["rates"], r["o_rates"])           # every accept / reject decision, by type
    scale = np.abs(r["Ho"]).max()
    assert np.abs(r["Href"] - r["Ho"]).max() <= 1e-13 * scale
    assert np.abs(r["Rref"] - r["Ro"]).max() <= 1e-13 * max(scale, np.abs(r["Ro"]).max())
    assert r["vol"] == pytest.approx(ho.volume(r["Ho"]), rel=1e-13)
    if not (r["props"]["type"] == ho.STRETCH).any():
        assert np.array_equal(r["Href"], r["Ho"]) and np.array_equal(r["Rref"], r["Ro"])


@pytest.mark.parametrize("shape,seed,sweeps,lf", CASES)
def test_reference_mc_run_bit_exact_with_deterministic_exp(shape, seed, sweeps, lf):
    """Same run with the single substitution np.exp -> hd_exp: trajectories are bit-identical."""
    r = run_both(live(True), shape, seed, sweeps, lf, dshear=0.4, dstretch=0.08)
    assert np.array_equal(r["rates"], r["o_rates"])
    assert np.array_equal(r["Href"].view(np.uint64), r["Ho"].view(np.uint64))
    assert np.array_equal(r["Rref"].view(np.uint64), r["Ro"].view(np.uint64))
    assert r["vol"] == ho.volume(r["Ho"])
    assert set(r["dec"]["reason"]) >= {0, 1}                  # the stream exercised accepts and overlaps


def test_reference_rejection_reasons_are_exercised():
    """The comparison above must have covered every branch of MCNS.py:349-380: ceiling, weight, overlap,
    aspect ratio, angle -- checked on the oracle's reasons for the same streams."""
    seen = set()
    for shape, seed, sweeps, lf in CASES:
        r = run_both(live(True), shape, seed, sweeps, lf, dshear=0.4, dstretch=0.08)
        seen |= set(int(x) for x in r["dec"]["reason"])
    # strong shears to hit the shape limits
    r = run_both(live(True), "tiny", 11, 12, None, dshear=3.0, dstretch=0.5)
    seen |= set(int(x) for x in r["dec"]["reason"])
    assert np.array_equal(r["rates"], r["o_rates"]) and np.array_equal(r["Rref"], r["Ro"])
    assert {0, 1, 2, 4, 5} <= seen, seen


def test_volume_ceiling_edges():
    """MCNS.py:351: accept iff u < boltz AND (V_new - limit) < DBL_EPSILON (absolute!).  Volume proposals that land
    on, one ulp above / below and far from the ceiling; oracle's early reject (1e-9 relative) must never differ."""
    L = live(True)
    cfg, mr, H0, R0 = start_state("tiny1", 21, dv_max=5.0)
    V0 = ho.volume(H0)
    dvs = [0.0, 1e-13 * V0, -1e-13 * V0, 3e-10 * V0, 1.0000001e-9 * V0, 0.9999999e-9 * V0, 1e-6 * V0, -0.3 * V0, 0.2 * V0,
           np.spacing(V0), -np.spacing(V0), 2e-9 * V0]
    mps = ho.moves_per_sweep(cfg.nbeads, cfg.nchains)
    for limit in (V0, np.nextafter(V0, 0), np.nextafter(V0, np.inf), V0 * (1 + 5e-10)):
        props = np.zeros(mps * 2, dtype=ho.PROPOSAL_DTYPE)
        props["type"] = ho.TRANS   # fill with harmless zero translations, then the volume proposals
        props["u_acc"] = 0.5
        for k, dv in enumerate(dvs):
            props[k]["type"] = ho.VOL
            props[k]["p"][0] = dv
            props[k]["u_acc"] = 0.3 if k % 2 else 0.0
        L.setup(cfg.nchains, cfg.nbeads, H0, R0)
        vol, rates = L.mc_run(cfg, props, move_ratio=mr, volume_limit=limit)
        Ho, Ro = H0.copy(), R0.copy()
        dec = ho.replay(cfg, Ho, Ro, props, volume_limit=limit)
        acc_vol = dec["accepted"][props["type"] == ho.VOL].sum()
        assert rates[0] == acc_vol / len(dvs)
        assert np.array_equal(L.alk.H[0], Ho) and np.array_equal(L.alk.R[0], Ro)
        assert vol <= limit


def _cells(rng, n):
    """Random cells: near-cubic, sheared, squashed."""
    out = []
    for _ in range(n):
        kind = rng.integers(3)
        H = np.eye(3) * rng.uniform(3, 12)
        if kind == 0:
            H += rng.normal(scale=0.3, size=(3, 3))
        elif kind == 1:
            H += rng.normal(scale=2.5, size=(3, 3))
        else:
            H = H * rng.uniform(0.4, 1.6, size=(3, 1)) + rng.normal(scale=1.0, size=(3, 3))
        out.append(H)
    return out


def test_shape_predicates_fuzz_against_reference_functions():
    """ho_aspect_fails / ho_angle_fails (cbrt- and arccos-free forms used by oracle and GPU) against the reference's
    own predicates `min_aspect_ratio(ibox) < L` (MCNS.py:97-118,194) and `min_angle(ibox) < A pi/180`
    (MCNS.py:120-133,190-196) on >= 1e5 random cells, and on cells constructed within 1e-12 of each limit."""
    L = live(False)
    L.setup(1, 1, np.eye(3), np.zeros((1, 3)))
    ref, alk = L.mod, L.alk
    rng = np.random.default_rng(2026)
    n_near = [0, 0]
    n_dis = [0, 0]
    cells = _cells(rng, 100000)
    lim_ar, lim_ang = 0.8, 45.0
    cos_lim = math.cos(lim_ang * math.pi / 180)
    lim_rad = lim_ang * np.pi / 180
    for H in cells:
        alk.H[0][...] = H
        ar = ref.min_aspect_ratio(1)
        ang = ref.min_angle(1)
        a_ref, g_ref = bool(ar < lim_ar), bool(ang < lim_rad)
        a_o, g_o = bool(ho.aspect_fails(H, lim_ar)), bool(ho.angle_fails(H, cos_lim))
        # the two formulas round differently: they may only disagree within a few ulp of the limit
        if a_ref != a_o:
            n_dis[0] += 1
            assert abs(ar - lim_ar) <= 8 * np.spacing(lim_ar)
        if g_ref != g_o:
            n_dis[1] += 1
            assert abs(ang - lim_rad) <= 8 * np.spacing(lim_rad)
    assert n_dis == [0, 0], n_dis   # (random cells never come that close)
    # near the limits: scale one cell vector / rotate one so that the value sits at limit * (1 +- d)
    for d in (1e-12, 1e-11, 1e-9, 1e-6):
        for H in cells[:400]:
            alk.H[0][...] = H
            ar = ref.min_aspect_ratio(1)
            for sgn in (-1.0, 1.0):
                # aspect: pick the limit relative to the value instead of deforming the cell
                lim = ar / (1.0 + sgn * d)
                assert bool(ref.min_aspect_ratio(1) < lim) == bool(ho.aspect_fails(H, lim)) == (sgn < 0)
                n_near[0] += 1
                ang = ref.min_angle(1)
                if ang < 1e-3 or ang > 1.5:
                    continue
                lim_r = ang / (1.0 + sgn * d)
                assert bool(ang < lim_r) == bool(ho.angle_fails(H, math.cos(lim_r))) == (sgn < 0)
                n_near[1] += 1
    assert min(n_near) > 2000


def test_shear_and_stretch_steps_direct():
    """box_shear_step / box_stretch_step (MCNS.py:135-252) called directly: delta_H, boltz and the ORDER of the three
    rejection tests (aspect, angle, overlap -- MCNS.py:194-199,243-248), read off the call trace."""
    rng = np.random.default_rng(5)
    for det in (False, True):
        L = live(det)
        for trial in range(300):
            shape = ("tiny", "tiny1", "t3")[trial % 3]
            cfg, mr, H0, R0 = start_state(shape, 100 + trial % 7, presweeps=1)
            if trial % 4 == 0:   # a squashed start so that the shape limits trigger
                H0 = H0 * np.array([[1.0], [1.0], [rng.uniform(0.55, 0.9)]])
                H0[0, 1] += rng.uniform(0, 0.8) * H0[0, 0]
            kind = ho.SHEAR if trial % 2 == 0 else ho.STRETCH
            step = float(rng.choice([0.05, 0.5, 2.0]))
            p = np.zeros(1, dtype=ho.PROPOSAL_DTYPE)
            p["type"] = kind
            i, j = int(rng.integers(3)), int(rng.integers(3))
            z = rng.normal(size=2)
            L.setup(cfg.nchains, cfg.nbeads, H0, R0)
            L.alk.cfg_kw = dict(min_ar=0.8, min_ang=45.0)
            L.alk.F.append((i + 0.5) / 3)
            if kind == ho.SHEAR:
                p["idx0"] = i
                p["p"][0, :2] = step * z
                L.rnd.normals.extend([(step, step * z[0]), (step, step * z[1])])
            else:
                L.alk.F.append((j + 0.5) / 3)
                p["idx0"], p["idx1"] = i, ((j + 1) % 3 if j == i else j)
                p["p"][0, 0] = step * z[0]
                L.rnd.normals.append((step, step * z[0]))
            L.alk.tracing, L.alk.trace, L.alk._last_state = True, [], None
            fn = L.mod.box_shear_step if kind == ho.SHEAR else L.mod.box_stretch_step
            boltz, dH = fn(1, step, 0.8, 45.0)
            L.alk.tracing = False
            assert L.drained()
            Ho, Ro = H0.copy(), R0.copy()
            c = ho.make_cfg(cfg.nchains, cfg.nbeads, mr, min_ar=0.8, min_ang=45.0)
            dec = ho.replay(c, Ho, Ro, p, mode=ho.MODE_PROPOSE)[0]
            assert int(boltz) == int(dec["boltz"])
            calls = [t[0] for t in L.alk.trace]
            reason = int(dec["reason"])
            if reason == 4:    # aspect ratio failed: nothing else was evaluated
                assert "alkane_check_chain_overlap" not in calls
            elif reason in (0, 1):
                assert calls[-1] == "alkane_check_chain_overlap"
            # the cell the reference left behind = the oracle's trial cell
            if kind == ho.SHEAR or det:
                assert np.array_equal(L.alk.H[0], Ho) and np.array_equal(L.alk.R[0], Ro)
                assert np.array_equal(dH, Ho - H0) or np.abs(dH - (Ho - H0)).max() <= 2 * np.spacing(np.abs(H0).max())
            else:
                assert np.abs(L.alk.H[0] - Ho).max() <= 4 * np.spacing(np.abs(Ho).max())


def test_adjust_mc_steps_rule_matches_update_steps():
    """The halve / double rule of adjust_mc_steps (MCNS.py:687-716) with the reference's own function: MC_run is
    replaced by scripted acceptance rates, everything else runs as written."""
    from hans_b200.ns import STEP_MAX, STEP_MIN, update_steps
    L = live(False)
    ref = L.mod
    rng = np.random.default_rng(9)

    class Comm:
        def Get_size(self): return 2

        def Allreduce(self, src, dst, op=None):
            dst[...] = 2.0 * np.asarray(src)   # two ranks with the same rates (x2 /2 is exact)

    real = ref.MC_run
    try:
        for trial in range(200):
            mr = np.array([3, 48, 32, 16, 3, 3], dtype=float) * (rng.random(6) > 0.2)
            if mr.sum() == 0:
                continue
            rates = rng.choice([0.0, 0.1, 0.19999, 0.2, 0.3, 0.5, 0.50001, 0.9, 1.0], size=6)
            steps = dict(dv_max=float(rng.choice([1e-10, 3e-10, 2.0, 30.0, 50.0])), dr_max=float(rng.choice([1e-10, 0.7, 26.0, 50.0])),
                         dt_max=float(rng.choice([1e-10, 0.43, 7.0])), dh_max=float(rng.choice([1e-10, 0.4, 9.0])),
                         dshear=float(rng.choice([1e-5, 1.5e-5, 1.0, 80.0])), dstretch=float(rng.choice([1e-5, 0.3, 4.0])))
            cfg, _, H0, R0 = start_state("tiny", 3, presweeps=0)
            args = L.setup(cfg.nchains, cfg.nbeads, H0, R0, steps={k: steps[k] for k in ("dv_max", "dr_max", "dt_max", "dh_max")})
            args.update(min_angle=45.0, min_aspect_ratio=0.8)
            L.rnd.ints.append(0)   # mc_box = np.random.randint(nwalkers), MCNS.py:676

            def fake_mc_run(ns_data, sweeps, ratio, ibox, vol_max, **kw):
                i = int(np.argmax(ratio))
                out = np.zeros(6)
                out[i] = rates[i]
                return 1.0, out
            ref.MC_run = fake_mc_run
            avg, dshear, dstretch = ref.adjust_mc_steps(args, Comm(), mr, 100.0, walklength=10, lower_bound=0.2, upper_bound=0.5,
                                                        min_dstep=STEP_MIN, dv_max=STEP_MAX["dv_max"], dr_max=STEP_MAX["dr_max"],
                                                        dshear=steps["dshear"], dstretch=steps["dstretch"])
            got = dict(dv_max=L.alk.alkane_get_dv_max(), dr_max=L.alk.alkane_get_dr_max(), dt_max=L.alk.alkane_get_dt_max(),
                       dh_max=L.alk.alkane_get_dh_max(), dshear=dshear, dstretch=dstretch)
            want = update_steps(steps, np.where(mr != 0, rates, 0.0), mr)
            assert got == want, (trial, mr, rates, steps, got, want)
            assert np.array_equal(avg, np.where(mr != 0, rates, 0.0))
    finally:
        ref.MC_run = real


def test_default_move_ratio_and_initial_cell_match():
    """default_move_ratio (MCNS.py:610-626) and the initial cubic cell (MCNS.py:628-631) of the reference against the
    mirror in hans_b200.MCNS / the oracle helper."""
    from hans_b200 import MCNS as mine
    ref = live(False).mod
    for nchains, nbeads in [(64, 1), (16, 6), (32, 12), (128, 6), (32, 2), (11, 3), (5, 4)]:
        d = dict(nchains=nchains, nbeads=nbeads, from_restart=0)
        assert np.array_equal(ref.default_move_ratio(d), mine.default_move_ratio(d))
        d["move_ratio"] = "3,48,32,0,3,3"
        assert np.array_equal(ref.default_move_ratio(d), mine.default_move_ratio(d))
        d2 = dict(d, from_restart=1, move_ratio=np.array([1.0, 2, 3, 4, 5, 6]))
        assert np.array_equal(ref.default_move_ratio(d2), mine.default_move_ratio(d2))
        side = 0.999 * np.eye(3) * np.cbrt(nbeads * nchains * 15)
        assert np.array_equal(ho.initial_cell(nbeads, nchains), side)
    with pytest.raises(Exception):
        mine.default_move_ratio(dict(nchains=4, nbeads=2, move_ratio="1,2,3"))
    with pytest.raises(Exception):
        ref.default_move_ratio(dict(nchains=4, nbeads=2, move_ratio="1,2,3", from_restart=0))
    # the README input has no from_restart key: the reference raises KeyError, the mirror treats it as 0
    assert np.array_equal(mine.default_move_ratio(dict(nchains=16, nbeads=6, move_ratio="3,48,32,0,3,3")),
                          [3, 48, 32, 0, 3, 3])


def test_reference_walker_copy_helpers():
    """mk_ase_config / import_ase_to_ibox / clone_walker of the reference (MCNS.py:30-53,510-524,550-580) on the
    stand-in: pins the scaling convention and the bead order chain-major that the restart / clone code relies on."""
    L = live(False)
    cfg, mr, H0, R0 = start_state("tiny", 4)
    H = np.stack([H0, H0 * 1.1])
    R = np.stack([R0, R0 * 1.1])
    L.setup(cfg.nchains, cfg.nbeads, H, R)
    ref = L.mod
    at = ref.mk_ase_config(1, cfg.nbeads, cfg.nchains, scaling=2.0)
    assert np.array_equal(at.positions, R0 * 2.0) and np.array_equal(at.cell, H0 * 2.0)
    ref.import_ase_to_ibox(at, 2, dict(nbeads=cfg.nbeads, nchains=cfg.nchains), scaling=0.5)
    assert np.array_equal(L.alk.H[1], H0) and np.array_equal(L.alk.R[1], R0)
    L.alk.R[1] += 1.0
    ref.clone_walker(2, 1)
    assert np.array_equal(L.alk.R[0], R0 + 1.0) and np.array_equal(L.alk.H[0], H0)


def test_committed_golden_vectors_are_current(tmp_path):
    """tests/golden/ref_live_* must be what the generator produces from the reference as it lies under /root/reference."""
    import json
    import os
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    gen = os.path.join(here, "golden", "make_ref_live_golden.py")
    old_npz = dict(np.load(os.path.join(here, "golden", "ref_live_replay.npz")))
    old_tr = json.load(open(os.path.join(here, "golden", "ref_live_trace.json")))
    src = open(gen).read().replace("HERE = os.path.dirname(os.path.abspath(__file__))", f"HERE = {str(tmp_path)!r}", 1)
    src = src.replace("ROOT = os.path.dirname(os.path.dirname(HERE))", f"ROOT = {os.path.dirname(here)!r}", 1)
    script = tmp_path / "gen.py"
    script.write_text(src)
    subprocess.run([sys.executable, str(script)], check=True, capture_output=True, cwd=os.path.dirname(here))
    new_npz = dict(np.load(tmp_path / "ref_live_replay.npz"))
    assert sorted(new_npz) == sorted(old_npz)
    for k in old_npz:
        assert np.array_equal(old_npz[k], new_npz[k]), k
    assert json.load(open(tmp_path / "ref_live_trace.json")) == old_tr


def test_reference_nematic_functions_match_the_numpy_restatement():
    """/root/reference/nematic.py imported unmodified (ase.io and matplotlib replaced by empty stand-ins; the Atoms stand-in
    supplies what ase would: scaled positions wrapped into [0, 1), minimum-image end-to-end vectors, the cell volume): its
    nematic_order_param (nematic.py:83-137) and trans_order_param (nematic.py:30-55; runs for Nbeads = 1 only, see
    oracle/order_params.py) against the restatement the GPU kernel is checked with (hans_order_params, tests/test_gpu_ns.py)."""
    import importlib.util
    import os
    import types
    from oracle import order_params as op

    class Cell(np.ndarray):
        def reciprocal(self):
            return np.linalg.inv(np.asarray(self)).T

    class At:
        def __init__(self, cell, pos):
            self.cell = np.asarray(cell).view(Cell)
            self.pos = np.asarray(pos)

        def get_volume(self): return float(abs(np.linalg.det(np.asarray(self.cell))))
        def get_positions(self): return self.pos.copy()
        def get_scaled_positions(self): return (self.pos @ np.linalg.inv(np.asarray(self.cell))) % 1.0

        def get_distance(self, a0, a1, mic=False, vector=False):
            d = self.pos[a1] - self.pos[a0]
            if mic:   # nearest image (what ase's find_mic returns): wrap, then brute force over the 27 neighbouring cells
                c = np.asarray(self.cell)
                f = d @ np.linalg.inv(c)
                d = (f - np.rint(f)) @ c
                cands = [d + np.array([i, j, k]) @ c for i in (-1, 0, 1) for j in (-1, 0, 1) for k in (-1, 0, 1)]
                d = min(cands, key=lambda v: float(v @ v))
            return d if vector else float(np.linalg.norm(d))

    saved = {k: sys.modules.get(k) for k in ("ase", "ase.io", "matplotlib", "matplotlib.pyplot")}
    try:
        ase = types.ModuleType("ase"); ase.io = types.ModuleType("ase.io")
        mpl = types.ModuleType("matplotlib"); mpl.pyplot = types.ModuleType("matplotlib.pyplot")
        sys.modules.update({"ase": ase, "ase.io": ase.io, "matplotlib": mpl, "matplotlib.pyplot": mpl.pyplot})
        spec = importlib.util.spec_from_file_location("_hans_reference_nematic", os.path.join(ref_live.REFERENCE_ROOT, "nematic.py"))
        nem = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(nem)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    rng = np.random.default_rng(4)
    for shape, seed in (("C1", 3), ("dimer", 4), ("p7", 5), ("C2", 6)):
        nchains, nbeads, mr = SHAPES[shape]
        cfg = ho.make_cfg(nchains, nbeads, mr)
        H, R = ho.grow_walkers(cfg, 4, seed)
        for w in range(4):   # sheared cells, beads pushed out of the cell by lattice vectors
            ho.change_box(cfg, H[w], R[w], rng.normal(scale=0.6, size=(3, 3)))
            R[w] += rng.integers(-2, 3, size=(R.shape[1], 1)) * H[w][rng.integers(3)]
        configs = [At(H[w], R[w]) for w in range(4)]
        if nbeads > 1:
            want = nem.nematic_order_param(configs, nbeads, nchains)
            got = op.nematic_order_param(H, R, nbeads, nchains)
            assert np.allclose(got, np.real(want), rtol=0, atol=1e-12), (shape, got, want)
        else:
            want = nem.trans_order_param(configs, nbeads, nchains)
            got = op.trans_order_param(H, R, nbeads, nchains)
            assert np.allclose(got, want, rtol=0, atol=1e-12), (shape, got, want)
