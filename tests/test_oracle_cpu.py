"""CPU tests of the oracle (oracle/): known-answer vectors, the reference's Python-specified
pieces restated in numpy, physical invariants, and the golden fixtures under tests/golden/."""
import json
import math
import os

import numpy as np
import pytest

from common import SHAPES, compress, make_walkers

HERE = os.path.dirname(os.path.abspath(__file__))


# ---- Philox4x32-10: Random123 known-answer vectors ---------------------------------------------------
KAT = [
    ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
    ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
    ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
     [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
]


@pytest.mark.parametrize("ctr,key,want", KAT)
def test_philox_known_answers(oracle, ctr, key, want):
    assert list(oracle.philox(ctr, key)) == want
    from hans_b200 import philox
    assert list(philox.philox4x32_10(ctr, key)) == want


# ---- deterministic math vs libm -------------------------------------------------------------------------
def test_detmath_accuracy(oracle):
    rng = np.random.default_rng(0)

    def ulps(a, b):
        return abs(a - b) / np.spacing(abs(b))
    assert max(ulps(oracle.det_exp(x), math.exp(x)) for x in rng.uniform(-60, 60, 4000)) <= 2
    assert max(ulps(oracle.det_log(x), math.log(x)) for x in np.exp(rng.uniform(-40, 40, 4000))) <= 3
    assert max(ulps(oracle.det_cbrt(x), float(np.cbrt(x))) for x in np.exp(rng.uniform(-40, 40, 4000))) <= 2
    for x in rng.uniform(-12, 12, 4000):
        s, c = oracle.det_sincos(x)
        assert abs(s - math.sin(x)) < 3e-16 and abs(c - math.cos(x)) < 3e-16
    for x in [0.5, 1.5, 2.5, -0.5, -1.5, -2.5, 0.49999999999999994, 1e10 + 0.5, -7.5, 3.2, 0.0]:
        assert oracle.det_rint(x) == np.rint(x)
    assert oracle.det_exp(0.0) == 1.0 and oracle.det_log(1.0) == 0.0 and oracle.det_cbrt(27.0) == 3.0


# ---- the pieces the reference spells out in Python, restated independently in numpy -------------------
def ref_min_aspect_ratio(cell):      # MCNS.py:97-118
    vol = abs(np.linalg.det(cell))
    best = np.inf
    for i in range(3):
        n = np.cross(cell[(i + 1) % 3], cell[(i + 2) % 3])
        n = n / np.sqrt(n @ n)
        best = min(best, abs(n @ cell[i]))
    return best / np.cbrt(vol)


def ref_min_angle(cell):             # MCNS.py:120-133
    best = np.inf
    for i in range(3):
        a, b = cell[(i + 1) % 3], cell[(i + 2) % 3]
        d = abs(a @ b / np.sqrt((a @ a) * (b @ b)))
        best = min(best, np.arccos(d))
    return best


def test_cell_predicates_match_reference_python(oracle):
    rng = np.random.default_rng(2)
    for _ in range(200):
        H = np.eye(3) * rng.uniform(3, 12) + rng.normal(scale=1.5, size=(3, 3))
        assert oracle.volume(H) == pytest.approx(abs(np.linalg.det(H)), rel=1e-13)
        assert oracle.min_aspect_ratio(H) == pytest.approx(ref_min_aspect_ratio(H), rel=1e-12)
        assert math.acos(oracle.max_abs_cos(H)) == pytest.approx(ref_min_angle(H), rel=1e-10, abs=1e-12)
        assert np.allclose(oracle.hinv(H) @ H, np.eye(3), atol=1e-12)
    # cube: aspect ratio 1, all angles 90 degrees
    assert oracle.min_aspect_ratio(np.eye(3) * 4) == pytest.approx(1.0, rel=1e-15)
    assert oracle.max_abs_cos(np.eye(3) * 4) == 0.0


def test_moves_per_sweep_and_sizes(oracle):
    # MCNS.py:296-301 and the size table of SURVEY.md section 8
    assert oracle.moves_per_sweep(1, 64) == 71
    assert oracle.moves_per_sweep(2, 32) == 71
    assert oracle.moves_per_sweep(3, 10) == 27
    assert oracle.moves_per_sweep(6, 16) == 87
    assert oracle.moves_per_sweep(12, 32) == 359
    assert oracle.moves_per_sweep(6, 128) == 647


def test_minimum_image_single_vs_27_images(oracle):
    """[A1]: the single-image rule gives the true minimum whenever it matters for the overlap test
    (r < sigma) and every perpendicular width exceeds 2 sigma."""
    rng = np.random.default_rng(7)
    n_close = 0
    for _ in range(3000):
        H = np.eye(3) * rng.uniform(2.6, 6) + rng.normal(scale=0.5, size=(3, 3))
        if ref_min_aspect_ratio(H) * np.cbrt(abs(np.linalg.det(H))) <= 2.0:
            continue
        ri = rng.uniform(-10, 10, 3)
        d = rng.normal(size=3)
        rj = ri + d / np.linalg.norm(d) * rng.uniform(0.2, 1.6) + rng.integers(-3, 4, 3) @ H
        r1, r27 = oracle.pair_r2(H, ri, rj), oracle.pair_r2_27(H, ri, rj)
        if r27 < 1.0 or r1 < 1.0:
            n_close += 1
            assert r1 == pytest.approx(r27, rel=1e-12)
    assert n_close > 500


# ---- move semantics (replay) -----------------------------------------------------------------------------
def _props(oracle, n=1):
    return np.zeros(n, dtype=oracle.PROPOSAL_DTYPE)


def test_translate_accept_and_exact_restore(oracle):
    cfg, H, R, _ = make_walkers("C1", 1, seed=3)
    H, R = H[0], R[0]
    R0 = R.copy()
    p = _props(oracle)
    p["type"], p["ichain"] = oracle.TRANS, 4
    p["p"][0, :3] = [0.01, -0.02, 0.005]
    d = oracle.replay(cfg, H, R, p)
    assert d["accepted"][0] == 1 and d["boltz"][0] == 1.0
    assert np.allclose(R[24:30], R0[24:30] + [0.01, -0.02, 0.005]) and np.array_equal(R[:24], R0[:24])
    # move chain 4 on top of chain 5: rejected, bit-exact restore (MCNS.py:379-380)
    R1 = R.copy()
    p["p"][0, :3] = R[30] - R[24] + 0.05
    d = oracle.replay(cfg, H, R, p)
    assert d["accepted"][0] == 0 and d["reason"][0] == 1 and np.array_equal(R, R1)
    # a lattice translation never changes the decision (periodic images)
    p["p"][0, :3] = 0.01 + H[0] - 2 * H[2]
    assert oracle.replay(cfg, H, R.copy(), p)["accepted"][0] == 1


def test_rotation_and_dihedral_preserve_geometry(oracle):
    cfg, H, R, _ = make_walkers("C4s", 1, seed=4)
    H, R = H[0], R[0]
    nb = 12
    props = oracle.gen_proposals(oracle.make_cfg(8, 12, [0, 0, 1, 1, 0, 0]), 5, 0, 0, 300)
    com0 = R.reshape(8, nb, 3).mean(1)
    dec = oracle.replay(cfg, H, R, props)
    assert dec["accepted"].sum() > 50
    x = R.reshape(8, nb, 3)
    b = np.diff(x, axis=1)
    assert np.allclose(np.linalg.norm(b, axis=2), 0.4, atol=1e-12)              # bond length
    cosang = -(b[:, 1:] * b[:, :-1]).sum(2) / 0.16
    assert np.allclose(np.degrees(np.arccos(cosang)), 109.7, atol=1e-9)          # bond angle
    rot_only = props["type"] == oracle.ROT
    assert rot_only.any() and (props["type"] == oracle.DIH).any()
    assert oracle.check_overlap(cfg, H, R) == 0
    # rotations alone keep the centroid
    R2 = make_walkers("C4s", 1, seed=4)[2][0]
    oracle.replay(cfg, H, R2, props[rot_only])
    assert np.allclose(R2.reshape(8, nb, 3).mean(1), com0, atol=1e-12)


def test_volume_move_rules(oracle):
    cfg, H, R, _ = make_walkers("C2", 1, seed=5)
    H, R = H[0], R[0]
    V0 = oracle.volume(H)
    p = _props(oracle)
    p["type"] = oracle.VOL
    # growth above the ceiling is rejected whatever u_acc is (MCNS.py:351)
    p["p"][0, 0], p["u_acc"] = 5.0, 0.0
    H1, R1 = H.copy(), R.copy()
    d = oracle.replay(cfg, H1, R1, p, volume_limit=V0)
    assert d["accepted"][0] == 0 and d["reason"][0] == 2 and np.array_equal(H1, H) and np.array_equal(R1, R)
    # V_new == limit is allowed (difference < epsilon)
    d = oracle.replay(cfg, H1, R1, p, volume_limit=V0 + 5.0 + 1e-9)
    assert d["accepted"][0] == 1 and oracle.volume(H1) == pytest.approx(V0 + 5.0, rel=1e-13)
    assert d["boltz"][0] == pytest.approx(((V0 + 5.0) / V0) ** 64, rel=1e-10)       # (V'/V)^nchains, MCNS.py:266
    # isotropic: the cell shape is unchanged, spheres keep their fractional coordinates
    assert np.allclose(H1 / np.cbrt((V0 + 5.0) / V0), H, rtol=1e-13)
    assert np.allclose(R1 @ np.linalg.inv(H1), R @ np.linalg.inv(H), atol=1e-12)
    # shrink: acceptance weight (V'/V)^n < 1 is compared with u_acc
    p["p"][0, 0] = -3.0
    w = ((V0 - 3.0) / V0) ** 64
    for u, acc in ((w * 0.999, 1), (w * 1.001, 0)):
        p["u_acc"] = u
        assert oracle.replay(cfg, H.copy(), R.copy(), p, volume_limit=1e300)["accepted"][0] == acc
    # pressure enters as exp(-P dV) (MCNS.py:266)
    cfgP = oracle.make_cfg(64, 1, [1, 192, 0, 0, 3, 3], pressure=0.7)
    p["u_acc"] = 0.0
    d = oracle.replay(cfgP, H.copy(), R.copy(), p, volume_limit=1e300)
    assert d["boltz"][0] == pytest.approx(math.exp(0.7 * 3.0) * w, rel=1e-10)


def test_shear_and_stretch_preserve_volume_and_revert(oracle):
    cfg, H, R, _ = make_walkers("C1", 1, seed=6)
    H, R = H[0], R[0]
    V0 = oracle.volume(H)
    gen = oracle.make_cfg(16, 6, [0, 0, 0, 0, 1, 1], dshear=0.3, dstretch=0.05)
    props = oracle.gen_proposals(gen, 9, 0, 0, 200)
    assert set(props["type"]) == {oracle.SHEAR, oracle.STRETCH}
    dec = oracle.replay(cfg, H, R, props)
    assert 0 < dec["accepted"].mean() < 1
    assert oracle.volume(H) == pytest.approx(V0, rel=1e-12)                       # MCNS.py:156-182, 231-232
    assert oracle.min_aspect_ratio(H) >= 0.8 and oracle.max_abs_cos(H) <= math.cos(math.radians(45.0))
    assert set(dec["reason"]) <= {0, 1, 4, 5}
    assert oracle.check_overlap(cfg, H, R) == 0
    # rejection reverts by -dH: equal to the start up to round-off, NOT bit-exact (MCNS.py:366-368)
    Hs, Rs = H.copy(), R.copy()
    p = _props(oracle)
    p["type"], p["idx0"], p["p"][0, 0], p["p"][0, 1] = oracle.SHEAR, 1, 40.0, 0.0   # violates the shape limits
    d = oracle.replay(cfg, Hs, Rs, p)
    assert d["accepted"][0] == 0 and d["reason"][0] in (4, 5)
    assert np.allclose(Hs, H, rtol=0, atol=1e-12) and np.allclose(Rs, R, rtol=0, atol=1e-10)


def test_generated_proposals_follow_documented_laws(oracle):
    cfg = oracle.make_cfg(16, 6, [3, 48, 32, 16, 3, 3], dr_max=0.7, dt_max=0.3, dh_max=0.2, dv_max=1.5,
                          dshear=0.25, dstretch=0.1)
    P = oracle.gen_proposals(cfg, 123, 4, 0, 40000)
    frac = np.bincount(P["type"], minlength=6) / len(P)
    assert np.allclose(frac, np.array([3, 48, 32, 16, 3, 3]) / 105, atol=0.01)     # MCNS.py:294,313-344
    assert P["ichain"].min() == 0 and P["ichain"].max() == 15
    t = P[P["type"] == oracle.TRANS]["p"][:, :3]
    assert np.abs(t).max() <= 0.7 and abs(t.mean()) < 0.01 and t.std() == pytest.approx(0.7 / math.sqrt(3), rel=0.03)
    q = P[P["type"] == oracle.ROT]["p"]
    assert np.allclose((q ** 2).sum(1), 1.0, atol=1e-14) and (2 * np.arccos(q[:, 0])).max() <= 0.3 + 1e-12
    d = P[P["type"] == oracle.DIH]
    assert set(d["idx0"]) == {3, 4, 5} and set(d["idx1"]) == {0, 1} and np.abs(d["p"][:, 0]).max() <= 0.2
    s = P[P["type"] == oracle.SHEAR]
    assert set(s["idx0"]) == {0, 1, 2} and s["p"][:, :2].std() == pytest.approx(0.25, rel=0.06)   # N(0, dshear)
    e = P[P["type"] == oracle.STRETCH]
    assert (e["idx0"] != e["idx1"]).all() and e["p"][:, 0].std() == pytest.approx(0.1, rel=0.08)
    assert 0.0 <= P["u_acc"].min() and P["u_acc"].max() < 1.0 and P["u_acc"].mean() == pytest.approx(0.5, abs=0.01)
    # counter-based: same (seed, walker, counter) -> same proposal; different walker -> different stream
    assert np.array_equal(oracle.gen_proposals(cfg, 123, 4, 100, 5), P[100:105])
    assert not np.array_equal(oracle.gen_proposals(cfg, 123, 5, 100, 5)["u_acc"], P[100:105]["u_acc"])


def test_growth_produces_valid_chains(oracle):
    for shape in ("C1", "C2", "C4s"):
        nchains, nbeads, mr = SHAPES[shape]
        cfg = oracle.make_cfg(nchains, nbeads, mr)
        H, R = oracle.grow_walkers(cfg, 3, seed=12)
        assert np.allclose(H[0], 0.999 * np.eye(3) * np.cbrt(15 * nchains * nbeads))   # MCNS.py:629
        for w in range(3):
            assert oracle.check_overlap(cfg, H[w], R[w]) == 0
            if nbeads > 1:
                b = np.diff(R[w].reshape(nchains, nbeads, 3), axis=1)
                assert np.allclose(np.linalg.norm(b, axis=2), 0.4, atol=1e-13)


# ---- physical known-answer test: hard-sphere equation of state ----------------------------------------
def test_hard_sphere_npt_equation_of_state(oracle):
    """External KAT: NPT sampling with the engine's own moves and the acceptance weight
    exp(-P dV + N ln V'/V) must reproduce the Carnahan-Starling fluid equation of state."""
    N, P = 64, 1.5
    cfg = oracle.make_cfg(N, 1, [2, 64, 0, 0, 0, 0], pressure=P, dr_max=0.35, dv_max=12.0)
    H, R = oracle.grow_walkers(cfg, 8, seed=21)
    H, R = compress(cfg, H, R, 0.35, 21)
    ctrs = np.zeros(8, dtype=np.uint64)
    ids = np.arange(8)
    oracle.mc_run_many(cfg, H, R, ids, 1500, 1e300, 33, ids, ctrs)      # equilibrate
    rho = []
    for _ in range(60):
        vols, _, _ = oracle.mc_run_many(cfg, H, R, ids, 60, 1e300, 33, ids, ctrs)
        rho.append(N / vols)
    rho = np.array(rho)
    eta = np.pi / 6 * rho.mean()
    # Carnahan-Starling: beta P sigma^3 = (6/pi) eta (1 + eta + eta^2 - eta^3) / (1 - eta)^3
    lo, hi = 0.01, 0.6
    for _ in range(60):
        mid = 0.5 * (lo + hi)
        if 6 / np.pi * mid * (1 + mid + mid ** 2 - mid ** 3) / (1 - mid) ** 3 < P:
            lo = mid
        else:
            hi = mid
    err = rho.mean(0).std(ddof=1) / np.sqrt(8) * np.pi / 6
    assert abs(eta - lo) < max(5 * err, 0.01), (eta, lo, err)


# ---- golden fixtures -----------------------------------------------------------------------------------------
def test_ns_analyse_port_matches_reference_output():
    """tests/golden/ns_analyse_golden.json holds what the reference's own ns_analyse printed
    (generated by tests/golden/make_ns_golden.py in the build container)."""
    from oracle import ns_analyse_port as nap
    g = json.load(open(os.path.join(HERE, "golden", "ns_analyse_golden.json")))
    lines = open(os.path.join(HERE, "golden", g["file"])).read().splitlines()
    K, n_cull, dof, flat, natoms = lines[0].split()
    assert (n_cull, flat) == ("1", "False") and lines[0].endswith(" ")
    vols = np.array([float(l.split()[1]) for l in lines[1:]])
    for r in g["rows"]:
        a = nap.analyse(vols, int(K), int(dof), r["T"])
        for k in ("V", "U", "Cvp", "log_Z"):
            assert a[k] == pytest.approx(r[k], rel=2e-5), (r["T"], k)   # the reference prints 6 significant digits
        assert a["mode_i"] == r["mode_i"]


def test_oracle_ns_reproduces_golden_curve(oracle):
    """A fresh oracle NS run from an unseen seed must fall inside the golden band."""
    from oracle import ns_oracle
    g = json.load(open(os.path.join(HERE, "golden", "ns_curves.json")))["B_tetramers8"]
    s = g["config"]
    r = ns_oracle.run_ns(s["nchains"], s["nbeads"], s["move_ratio"], s["K"], s["walklength"], s["iterations"], seed=4242)
    lnv = np.log(r["volumes"])[g["checkpoints"]]
    z = (lnv - np.array(g["mean_lnV"])) / np.maximum(np.array(g["sd_lnV"]), 1e-3)
    assert np.abs(z).max() < 5.0, z
    # not strictly monotone: a rejected shear/stretch reverts by -dH (MCNS.py:366-368), which may leave the
    # volume a few ulp above the ceiling it was sampled under
    assert (np.diff(r["volumes"]) <= 1e-9).all()
    # packing fraction of the final configuration via the fused-sphere chain volume
    # (Intersecting_Sphere_Notes.ipynb): V_chain(4) = 4/3 pi r^3 + pi (n-1)(16 r^3 - (d-2r)^2 (d+4r))/12
    rr, d, n = 0.5, 0.4, 4
    vchain = 4 / 3 * np.pi * rr ** 3 + np.pi * (n - 1) * (16 * rr ** 3 - (d - 2 * rr) ** 2 * (d + 4 * rr)) / 12
    assert 0.0 < 8 * vchain / r["volumes"][-1] < 0.74


# ---- the numpy restatement of nematic.py on configurations with known answers -------------------
def test_order_parameter_known_answers():
    from oracle import order_params as op
    cell = np.array([[6.0, 0.0, 0.0], [1.0, 5.0, 0.0], [0.5, -0.7, 7.0]])
    nb, nc = 4, 6
    rng = np.random.default_rng(3)
    # all chains along one direction: P2 = 1
    u = np.array([0.0, 0.6, 0.8])
    pos = np.concatenate([rng.uniform(0, 5, 3) + 0.4 * np.arange(nb)[:, None] * u for _ in range(nc)])
    assert abs(op.nematic_order_param([cell], [pos], nb, nc)[0] - 1.0) < 1e-12
    # chains along x, y, z in equal numbers: Q = 0, P2 = 0
    axes = np.eye(3)
    pos = np.concatenate([rng.uniform(0, 5, 3) + 0.4 * np.arange(nb)[:, None] * axes[c % 3] for c in range(nc)])
    assert abs(op.nematic_order_param([cell], [pos], nb, nc)[0]) < 1e-12
    # an end-to-end vector across the periodic boundary is taken at its minimum image
    far = pos.copy()
    far[nb - 1] += cell[0]
    assert np.allclose(op.nematic_order_param([cell], [far], nb, nc), op.nematic_order_param([cell], [pos], nb, nc))
    # single spheres on lattice points with integer fractional sums: translational order 1; random: ~ 1/sqrt(N)
    frac = np.array([[i, j, k] for i in range(3) for j in range(3) for k in range(3)]) / 1.0
    assert abs(op.trans_order_param([cell], [frac @ cell], 1, 27)[0] - 1.0) < 1e-9
    frac = rng.uniform(0, 1, (4000, 3))
    assert op.trans_order_param([cell], [frac @ cell], 1, 4000)[0] < 0.06
    assert op.nematic_order_param([cell], [frac[:27] @ cell], 1, 27)[0] == 0.0


# ---- the oracle against vectors made by executing the unmodified reference (tests/golden/make_ref_live_golden.py) ----
@pytest.mark.parametrize("shape", ["tiny1", "tiny", "p7", "t3", "C1", "C2", "h15", "dimer", "C4s"])
def test_oracle_reproduces_reference_golden(oracle, shape):
    """tests/golden/ref_live_replay.npz holds what /root/reference/NesSa/MCNS.py:MC_run made of explicit proposals
    (run in the build container; the file travels, the reference does not).  Decisions must be identical; the state
    bit-identical to the run with the deterministic exp, and within 1e-13 of the run with numpy's exp."""
    z = np.load(os.path.join(HERE, "golden", "ref_live_replay.npz"))
    tr = json.load(open(os.path.join(HERE, "golden", "ref_live_trace.json")))
    nchains, nbeads, sweeps = (int(x) for x in z[shape + ".shape"])
    cfg = oracle.make_cfg(nchains, nbeads, z[shape + ".move_ratio"], min_ar=tr["min_ar"], min_ang=tr["min_ang"], **tr["step_kw"])
    props = z[shape + ".props"].view(oracle.PROPOSAL_DTYPE)
    assert len(props) == sweeps * oracle.moves_per_sweep(nbeads, nchains)
    H, R = z[shape + ".H0"].copy(), z[shape + ".R0"].copy()
    dec = oracle.replay(cfg, H, R, props, volume_limit=float(z[shape + ".limit"][0]))
    att = np.bincount(props["type"], minlength=6).astype(float)
    acc = np.bincount(props["type"], weights=dec["accepted"], minlength=6)
    rates = acc / np.where(att == 0, 1, att)
    assert np.array_equal(rates, z[shape + ".rates_np"]) and np.array_equal(rates, z[shape + ".rates_det"])
    assert np.array_equal(H.view(np.uint64), z[shape + ".H_det"].view(np.uint64))
    assert np.array_equal(R.view(np.uint64), z[shape + ".R_det"].view(np.uint64))
    assert oracle.volume(H) == float(z[shape + ".vol_det"][0])
    scale = max(np.abs(R).max(), np.abs(H).max())
    assert np.abs(R - z[shape + ".R_np"]).max() <= 1e-13 * scale and np.abs(H - z[shape + ".H_np"]).max() <= 1e-13 * scale


# ---- the arithmetic shortcuts the oracle shares with the kernels, against the reference's own formulas --------------------
def test_volume_move_weight_equals_reference_formula(oracle):
    """The acceptance weight of a volume move is evaluated as (V'/V)^nchains * exp(-P dV) by repeated multiplication
    (hd_powi); the reference's formula is exp(-P dV + nchains ln(V'/V)) (mirrored at /root/reference/NesSa/MCNS.py:266).
    They must agree to rounding (<= 1e-13 relative for up to 128 chains), so `u < boltz` (MCNS.py:351) can only differ for
    u within that distance of the weight."""
    rng = np.random.default_rng(12)
    worst = 0.0
    for nchains in (1, 2, 16, 32, 64, 128):
        for P in (0.0, 2e-3):
            cfg = oracle.make_cfg(nchains, 1, [1, 0, 0, 0, 0, 0], pressure=P)
            L = 6.0 * nchains ** (1 / 3)
            H0 = np.eye(3) * L
            R0 = (np.indices((6, 6, 6)).reshape(3, -1).T[:nchains] + 0.5) * (L / 6.0)   # simple cubic: no overlaps
            V0 = oracle.volume(H0)
            for _ in range(40):
                dV = float(rng.uniform(-0.2, 0.3) * V0)
                p = np.zeros(1, dtype=oracle.PROPOSAL_DTYPE)
                p["type"], p["p"][0, 0] = oracle.VOL, dV
                H, R = H0.copy(), R0.copy()
                dec = oracle.replay(cfg, H, R, p, mode=oracle.MODE_PROPOSE)[0]
                if dec["boltz"] == 0.0:      # (a shrink that made spheres overlap)
                    continue
                want = math.exp(-P * dV + nchains * math.log((V0 + dV) / V0))
                worst = max(worst, abs(dec["boltz"] - want) / want)
    assert 0 < worst < 1e-13, worst


def test_minimum_image_equals_numpy_statement(oracle):
    """ho_pair_r2 (H^-1 by cofactors times ONE reciprocal of the determinant, FMA chains) against the plain numpy statement of
    the same rule (np.linalg.inv, nearest-integer wrap, back through H) on random and sheared cells: equal to 1e-12, so the
    strict decision r^2 < 1 can only differ for pairs within that distance of contact."""
    rng = np.random.default_rng(13)
    worst = 0.0
    for _ in range(4000):
        H = np.eye(3) * rng.uniform(3, 15) + rng.normal(scale=1.2, size=(3, 3))
        if abs(np.linalg.det(H)) < 5:
            continue
        ri, rj = rng.uniform(-20, 20, size=3), rng.uniform(-20, 20, size=3)
        s = (rj - ri) @ np.linalg.inv(H)
        s -= np.rint(s)
        m = s @ H
        want = float(m @ m)
        got = oracle.pair_r2(H, ri, rj)
        worst = max(worst, abs(got - want) / max(want, 1e-3))
    assert worst < 1e-12, worst
