"""CPU: the plain-C restatement (oracle/hp_oracle.c) against golden vectors generated from the reference itself
(tests/golden/make_golden.py over oracle/_ref = the unmodified reference sources compiled here)."""
import numpy as np
import pytest

from cases import CASES
from common import golden, oracle_cfg, check_tree_against_golden, rel_inf


def test_tables_match_definitions(oracle):
    t = oracle.tables()
    # LegendreCoeffientCount incl. the reference's truncation at degree 6 (Utility.h:87-106)
    assert t["counts"].tolist() == [1, 4, 10, 20, 35, 56, 83, 120, 165, 220, 286, 364, 455]
    a = np.arange(13)[:, None]
    d = np.arange(11)[None, :]
    assert np.allclose(t["nl"], np.sqrt((2 * a + 1) * 2.0 ** d), rtol=4.5e-16, atol=0)    # Newton-iterated (Utility.h:25-35): within 1 ulp of sqrt
    # BasisIndexValues: shell by shell, i then j, k = p-i-j
    assert t["basis_idx"][:10].tolist() == [[0, 0, 0], [0, 0, 1], [0, 1, 0], [1, 0, 0], [0, 0, 2], [0, 1, 1], [0, 2, 0],
                                            [1, 0, 1], [1, 1, 0], [2, 0, 0]]
    assert (t["basis_idx"].sum(1)[[0, 1, 4, 10, 20, 35, 56, 84, 120]] == np.arange(9)).all()
    # Gauss-Legendre rules integrate polynomials exactly; weights sum to 2
    for n in (1, 2, 9, 13, 17, 25, 45, 64):
        o = n * (n - 1) // 2
        r, w = t["roots"][o:o + n], t["weights"][o:o + n]
        assert abs(w.sum() - 2.0) < 1e-14
        assert abs((w * r ** (2 * n - 2)).sum() - 2.0 / (2 * n - 1)) < 1e-14
        ref_r, ref_w = np.polynomial.legendre.leggauss(n)
        idx = np.argsort(r, kind="stable")
        assert np.abs(r[idx] - ref_r).max() < 1e-14 and np.abs(w[idx] - ref_w).max() < 1e-14


def test_single_fits_match_reference_golden(oracle):
    from oracle import hpref
    g = golden("fits")
    names = list(CASES)
    for k in range(int(g["n_fits"])):
        case, degree, depth, degree_in = [int(v) for v in g["fit%03d_meta" % k]]
        c = CASES[names[case]]
        cfg, prog = hpref.make_config(**c["cfg"]), hpref.make_program(c["prog"])
        cell = g["fit%03d_cell" % k].astype(np.float64)
        cin = g["fit%03d_cin" % k] if degree_in else None
        coeffs, err = oracle.oracle_fit(cfg, prog, cell[:3] - cell[3], cell[:3] + cell[3], degree, depth, degree_in, cin)
        assert rel_inf(coeffs, g["fit%03d_coeffs" % k]) <= 1e-13
        assert abs(err - float(g["fit%03d_err" % k])) <= 1e-12 * abs(float(g["fit%03d_err" % k])) + 1e-300


@pytest.mark.parametrize("name", ["c1_readme", "sphere_poly_1e8", "csg_cont", "custom_domain", "csg_small"])
def test_tree_matches_reference_golden(oracle, name):
    from oracle import hpref
    cfg, prog = oracle_cfg(hpref, name)
    t = oracle.OracleTree.build(cfg, prog, threads=8, cg_tol=1e-13)
    g = golden(name)
    blk = hpref.parse_block(t.block())
    tol = 1e-10 if CASES[name]["cfg"].get("continuity", True) else 1e-13     # CG iterate is not pinned, its limit is
    worst, ndiv = check_tree_against_golden(blk, g, oracle.NCOUNT, tol)
    assert ndiv == 0          # the restatement is bit-identical to the reference: no tie can resolve differently
    st = t.stats()
    assert st["applied_p"] == float(g["applied_p"]) and st["applied_h"] == float(g["applied_h"])
    assert abs(st["final_total"] - float(g["final_total"])) <= 1e-12 * abs(float(g["final_total"]))
    q = t.query(g["query_pts"])
    assert np.abs(q - g["query_vals"]).max() <= (1e-9 if tol > 1e-12 else 1e-13)
    # outside the root: DBL_MAX (Octree.cpp:668-671) — the golden points spill 2 % over the root on purpose
    assert (g["query_vals"] == np.finfo(np.float64).max).any()


def test_mc_counter_tree_matches_reference_golden(oracle):
    """mc_counter nearness (seed 2017): the fixture was built by the reference's sources calling their own FApprox 100 times
    per fit on Philox points (tests/golden/make_golden.py mc); the restatement reproduces it."""
    from oracle import hpref
    cfg, prog = oracle_cfg(hpref, "sphere_exp_1e8")
    t = oracle.OracleTree.build(cfg, prog, threads=8, mc_seed=2017)
    g = golden("sphere_exp_1e8_mc2017")
    worst, ndiv = check_tree_against_golden(hpref.parse_block(t.block()), g, oracle.NCOUNT, 1e-13)
    assert ndiv == 0
    st = t.stats()
    assert st["applied_p"] == float(g["applied_p"]) and st["applied_h"] == float(g["applied_h"])
    assert np.abs(t.query(g["query_pts"]) - g["query_vals"]).max() <= 1e-13
    # and it is not the exact-mean tree
    assert int(g["n_coeffs"]) != int(golden("sphere_poly_1e8")["n_coeffs"])


def test_memory_block_round_trip(oracle):
    from oracle import hpref
    cfg, prog = oracle_cfg(hpref, "csg_small")
    t = oracle.OracleTree.build(cfg, prog, threads=8)
    b = t.block()
    assert len(b) == 96 + 8 * hpref.parse_block(b)["n_coeffs"] + 56 * hpref.parse_block(b)["n_nodes"]
    t2 = oracle.OracleTree.from_block(b)
    pts = np.random.default_rng(3).uniform(-0.25, 0.5, (5000, 3))
    assert np.array_equal(t.query(pts), t2.query(pts))
    assert np.array_equal(np.asarray(t2.block()), np.asarray(b))
    with pytest.raises(ValueError):
        oracle.OracleTree.from_block(b[:-8])


def test_query_gradient_is_unit_and_matches_finite_difference(oracle):
    from oracle import hpref
    cfg, prog = oracle_cfg(hpref, "c1_readme")
    cfg.continuity_enforce = 0
    t = oracle.OracleTree.build(cfg, prog, threads=8)
    pts = np.random.default_rng(5).uniform(-0.2, 0.45, (500, 3))
    v, g = t.query_gradient(pts)
    assert np.allclose(np.linalg.norm(g, axis=1), 1.0, atol=1e-12)
    assert np.array_equal(v, t.query(pts))
    # sphere: gradient points away from the centre
    n = pts / np.linalg.norm(pts, axis=1)[:, None]
    assert (np.einsum("ij,ij->i", n, g) > 0.99).all()


def test_mesh_distance_matches_reference_golden(oracle):
    """Float32 mesh signed distance restatement (oracle/hp_oracle_mesh.h) against Mesh::SignedDistanceAtPt of the
    reference (golden generated through its own BVH): bit-exact."""
    from meshgen import bumpy_torus
    v, t = bumpy_torus(60, 40)
    m = oracle.OracleMesh(v, t)
    g = golden("mesh_torus")
    assert np.array_equal(m.sdf(g["pts"], True, 8), g["sdf"])
    assert np.array_equal(m.sdf(g["pts"][:1500], False, 8), g["sdf"][:1500])          # brute force == BVH (MeshingUnitTests.cpp:92-138)
    with pytest.raises(ValueError):
        oracle.OracleMesh(v, t[:-1])                                                  # open mesh: an edge has no twin
