"""CPU: pin the C restatement against the reference's own sources compiled here (oracle/_ref). Skipped where
oracle/_ref was not built. The parity builds use -ffp-contract=off and the restatement follows the reference's
operation order, so everything except the CG iterate is BIT-IDENTICAL."""
import numpy as np
import pytest

from cases import CASES
from common import oracle_cfg


def test_tables_bit_identical(ref, oracle):
    a, b = ref.ref_tables(), oracle.tables()
    for k in a:
        assert np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("degree,depth,degree_in", [(2, 4, 0), (3, 5, 0), (4, 4, 3), (6, 5, 0), (7, 4, 6)])
def test_fit_bit_identical(ref, oracle, degree, depth, degree_in):
    cfg, prog = oracle_cfg(ref, "c2_csg")
    half = 0.5 ** (depth + 1)
    centre = (np.array([5, 9, 7]) % 2 ** depth + 0.5) * 2 * half - 0.5
    cin = np.random.default_rng(1).normal(size=oracle.NCOUNT[degree_in]) if degree_in else None
    a, ea = ref.ref_fit(cfg, prog, centre - half, centre + half, degree, depth, degree_in, cin)
    b, eb = oracle.oracle_fit(cfg, prog, centre - half, centre + half, degree, depth, degree_in, cin)
    assert np.array_equal(a, b) and ea == eb


@pytest.mark.parametrize("name", ["sphere_exp_1e8", "csg_small", "custom_domain"])
def test_build_bit_identical(ref, oracle, name):
    cfg, prog = oracle_cfg(ref, name)
    r = ref.RefTree.build(cfg, prog, mode=1, threads=8)
    o = oracle.OracleTree.build(cfg, prog, threads=8)
    assert np.array_equal(r.apply_log(), o.apply_log())          # same jobs applied in the same order with the same errors
    br, bo = ref.parse_block(r.block()), ref.parse_block(o.block())
    assert np.array_equal(br["coeffs"], bo["coeffs"])
    for f in ("child", "mn", "mx", "deg", "depth"):
        assert np.array_equal(br["nodes"][f], bo["nodes"][f]), f
    pts = np.random.default_rng(0).uniform(-0.6, 0.6, (20000, 3)) * np.float64(cfg.root_max[0] - cfg.root_min[0]) + 0.5 * (cfg.root_max[0] + cfg.root_min[0])
    assert np.array_equal(r.query(pts, 8), o.query(pts, 8))
    # the reference's FromMemoryBlock + Query accepts the restatement's block and vice versa
    assert np.array_equal(ref.RefTree.from_block(o.block()).query(pts[:2000]), r.query(pts[:2000]))
    assert np.array_equal(oracle.OracleTree.from_block(r.block()).query(pts[:2000]), r.query(pts[:2000]))


@pytest.mark.parametrize("name", ["sphere_exp_1e8", "sphere_poly_1e8"])
def test_mc_counter_nearness_bit_identical(ref, oracle, name):
    """mc_counter(seed): the reference's 100-sample nearness estimator (Octree.cpp:1209-1247) through the reference's own
    FApprox on Philox points == the restatement, bit for bit; it is a different tree than the exact-mean one, and another
    seed gives another estimate."""
    cfg, prog = oracle_cfg(ref, name)
    r = ref.RefTree.build(cfg, prog, mode=1, threads=8, mc_seed=2017)
    o = oracle.OracleTree.build(cfg, prog, threads=8, mc_seed=2017)
    assert np.array_equal(r.apply_log(), o.apply_log())
    br, bo = ref.parse_block(r.block()), ref.parse_block(o.block())
    assert np.array_equal(br["coeffs"], bo["coeffs"])
    for f in ("child", "mn", "mx", "deg", "depth"):
        assert np.array_equal(br["nodes"][f], bo["nodes"][f]), f
    exact = oracle.OracleTree.build(cfg, prog, threads=8)
    other = oracle.OracleTree.build(cfg, prog, threads=8, mc_seed=2018)
    la, le, lo = o.apply_log(), exact.apply_log(), other.apply_log()
    assert la.shape != le.shape or not np.array_equal(la, le)
    assert la.shape != lo.shape or not np.array_equal(la, lo)
    # the estimate is close to the mean it estimates: the trees have similar sizes
    assert abs(len(la) - len(le)) <= 0.2 * len(le)


def test_max_degree_and_exact_total_switches(ref, oracle):
    cfg, prog = oracle_cfg(ref, "sphere_poly_1e8")
    for kw in (dict(max_degree=3), dict(total_mode=1), dict(max_degree=2, max_depth=6)):
        r = ref.RefTree.build(cfg, prog, mode=1, threads=8, **kw)
        o = oracle.OracleTree.build(cfg, prog, threads=8, **kw)
        assert np.array_equal(r.apply_log(), o.apply_log()), kw
        assert np.array_equal(np.asarray(r.block()[:8 + 8 * 100]), np.asarray(o.block()[:8 + 8 * 100]))


def test_continuity_matrix_and_converged_solution(ref, oracle):
    import scipy.sparse as sp
    cfg, prog = oracle_cfg(ref, "sphere_cont_1e8")
    r = ref.RefTree.build(cfg, prog, mode=1, threads=8, cg_tol=1e-13)
    o = oracle.OracleTree.build(cfg, prog, threads=8, cg_tol=1e-13)
    br, bo = ref.parse_block(r.block()), ref.parse_block(o.block())
    n = br["n_coeffs"]
    rows, cols, vals = r.continuity_triplets()
    m_ref = sp.coo_matrix((vals, (rows, cols)), shape=(n, n)).tocsr()
    m_ref.sum_duplicates()
    rp, col, val = o.continuity_csr(False)
    m_or = sp.csr_matrix((val, col, rp), shape=(n, n))
    assert m_ref.nnz == m_or.nnz
    assert abs(m_ref - m_or).max() <= 1e-12 * abs(m_ref).max()
    # converged solutions of (M + lambda I) x = lambda c agree; the CG iterate itself is unpinned (Eigen's IC absent)
    assert np.abs(br["coeffs"] - bo["coeffs"]).max() <= 1e-12 * np.abs(br["coeffs"]).max()


def test_literal_reference_create_agrees_within_its_own_test_tolerance(ref, oracle):
    """The shipped Create (async pool, rand() nearness) is non-deterministic; it and the deterministic
    schedule both pass the reference's own acceptance test: |Query - analytic| <= 0.01 (HPUnitTests.cpp:46-77)."""
    cfg, prog = oracle_cfg(ref, "sphere_poly_1e8")
    lit = ref.RefTree.build(cfg, prog, mode=0, threads=8)
    det = oracle.OracleTree.build(cfg, prog, threads=8)
    pts = np.random.default_rng(11).uniform(-0.5, 0.5, (100000, 3))
    truth = np.linalg.norm(pts - np.array([0.25, 0, 0]), axis=1) - 0.5
    assert np.abs(lit.query(pts, 8) - truth).max() <= 0.01
    assert np.abs(det.query(pts, 8) - truth).max() <= 0.01
    assert ref.parse_block(lit.block())["n_nodes"] == ref.parse_block(det.block())["n_nodes"]


def test_mesh_distance_bit_identical(ref, oracle):
    """Mesh::SignedDistanceAtPt through the reference's own BVH and by brute force == the restatement, bit for bit; also on
    the one mesh the reference ships (Resources/halfedge_fail.obj) where /root/reference is present."""
    import os
    from meshgen import bumpy_torus, mesh_root
    v, t = bumpy_torus(48, 32)
    rm, om = ref.RefMesh.create(v, t, bvh=True), oracle.OracleMesh(v, t)
    lo, hi = mesh_root(v)
    pts = np.random.default_rng(0).uniform(lo, hi, (20000, 3)).astype(np.float32)
    a = rm.sdf(pts, True, 8)
    assert np.array_equal(a, om.sdf(pts, True, 8))
    assert np.array_equal(rm.sdf(pts[:1000], False, 8), om.sdf(pts[:1000], False, 8))
    assert np.array_equal(a[:1000], rm.sdf(pts[:1000], False, 8))                     # the reference's own BVH-vs-brute-force test
    assert ref.RefMesh.create(v, t[:-1]) is None                                     # CreateHalfEdges fails on an open mesh
    obj = "/root/reference/Resources/halfedge_fail.obj"
    if os.path.exists(obj):
        rm2 = ref.RefMesh.load_obj(obj, bvh=True)
        v2, t2 = rm2.arrays()
        om2 = oracle.OracleMesh(v2, t2)
        lo2, hi2 = mesh_root(v2)
        p2 = np.random.default_rng(1).uniform(lo2, hi2, (4000, 3)).astype(np.float32)
        assert np.array_equal(rm2.sdf(p2, True, 8), om2.sdf(p2, True, 8))


def ray_batch(cfg_kwargs, n, seed):
    """Rays for QueryRay: origins inside and outside the root box, unit directions (some axis-aligned: 1/0 = inf in Ray::Ray)."""
    rng = np.random.default_rng(seed)
    mn = np.asarray(cfg_kwargs.get("root_min", (-0.5,) * 3), np.float64)
    mx = np.asarray(cfg_kwargs.get("root_max", (0.5,) * 3), np.float64)
    c, e = (mn + mx) / 2, (mx - mn)
    o = c + rng.uniform(-0.9, 0.9, (n, 3)) * e
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1)[:, None]
    d[: n // 20] = np.eye(3)[rng.integers(0, 3, n // 20)] * rng.choice([-1.0, 1.0], (n // 20, 1))
    aim = rng.random(n) < 0.5                         # half of them aimed at the middle of the box
    t = (c + rng.uniform(-0.2, 0.2, (n, 3)) * e) - o
    d[aim] = (t / np.linalg.norm(t, axis=1)[:, None])[aim]
    return o, d


@pytest.mark.parametrize("name", ["sphere_exp_1e8", "custom_domain"])
def test_query_ray_bit_identical(ref, oracle, name):
    """Octree::QueryRay (Octree.cpp:705-746): the restatement against the reference's own code, hit flags and t bit for bit —
    on the default root box and on a custom one, where the reference's double root mapping shows."""
    cfg, prog = oracle_cfg(ref, name)
    r = ref.RefTree.build(cfg, prog, mode=1, threads=8)
    o = oracle.OracleTree.build(cfg, prog, threads=8)
    org, d = ray_batch(CASES[name]["cfg"], 20000, 3)
    for t_max in (0.3, 10.0):
        hr, tr = r.query_ray(org, d, t_max)
        ho, to = o.query_ray(org, d, t_max)
        assert np.array_equal(hr, ho) and np.array_equal(tr, to)
        if name == "sphere_exp_1e8":
            assert 0 < hr.sum() < len(hr)          # (on the custom root box the reference's double mapping makes every ray miss)
