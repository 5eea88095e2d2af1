"""GPU: device mesh SDF (BVH + closest triangle + angle-weighted pseudonormals, float32) against the reference's
Mesh::SignedDistanceAtPt — bit-exact, because a 1-ulp float32 difference would already break the 1e-10 coefficient bar."""
import numpy as np
import pytest

from common import golden, rel_inf
from cases import leaf_table, path_code, divergent_cells
from meshgen import bumpy_torus, mesh_root

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torus(hp):
    v, t = bumpy_torus(60, 40)
    return v, t, hp.Mesh(v, t)


def test_mesh_distance_is_bit_identical_to_the_reference_golden(torus):
    v, t, m = torus
    g = golden("mesh_torus")
    d = m.SignedDistanceAtPt(g["pts"])
    assert np.array_equal(d, g["sdf"]), "mismatches: %d, max |d| %g" % ((d != g["sdf"]).sum(), np.abs(d - g["sdf"]).max())


def test_mesh_distance_matches_oracle_brute_force_and_bvh(hp, oracle, torus):
    v, t, m = torus
    om = oracle.OracleMesh(v, t)
    lo, hi = mesh_root(v)
    rng = np.random.default_rng(7)
    pts = rng.uniform(lo, hi, (50000, 3)).astype(np.float32)
    near = (v[rng.integers(0, len(v), 5000)] + rng.normal(0, 1e-3, (5000, 3))).astype(np.float32)     # hugging the surface
    onv = v[:2000]                                                                                     # exactly on vertices
    allp = np.concatenate([pts, near, onv])
    d = m.SignedDistanceAtPt(allp)
    assert np.array_equal(d, om.sdf(allp, True, 8))
    assert np.array_equal(d[:3000], om.sdf(allp[:3000], False, 8))            # brute force, Mesh.cpp:134-159
    mn, mx = m.CalculateMeshAABB()
    assert np.array_equal(mn, v.min(0)) and np.array_equal(mx, v.max(0))
    assert m.SignedDistanceAtPt(np.array([[2.0, 2.0, 2.0]]))[0] > 0 and m.SignedDistanceAtPt(np.array([[0.3, 0.0, 0.0]]))[0] < 0
    assert len(m.SignedDistanceAtPt(np.zeros((0, 3)))) == 0


def test_fine_mesh_far_and_near_points_match_the_oracle(hp, oracle):
    """139 k triangles of 3 mm against distances up to 0.4: the regime where the oriented-box pruning of mesh_eval.cuh does
    most of its work (hundreds of axis-aligned leaf boxes overlap the search sphere). Still bit-identical."""
    v, t = bumpy_torus(400, 174)
    m = hp.Mesh(v, t)
    om = oracle.OracleMesh(v, t)
    lo, hi = mesh_root(v)
    rng = np.random.default_rng(11)
    far = rng.uniform(lo, hi, (60000, 3)).astype(np.float32)
    axis = np.stack([rng.normal(0, 1e-3, 4000), rng.normal(0, 1e-3, 4000), rng.uniform(-0.3, 0.3, 4000)], 1).astype(np.float32)   # torus axis: many equidistant triangles
    cent = v[t[rng.integers(0, len(t), 30000)]].mean(1)
    near = (cent + rng.normal(0, 1, (30000, 3)) * rng.choice([1e-5, 1e-4, 1e-3, 1e-2], (30000, 1))).astype(np.float32)
    allp = np.concatenate([far, axis, near, v[:3000]])
    d = m.SignedDistanceAtPt(allp)
    ref = om.sdf(allp, True, 16)
    assert np.array_equal(d, ref), "mismatches: %d of %d, max |diff| %g" % ((d != ref).sum(), len(d), np.abs(d - ref).max())


def test_open_mesh_is_rejected(hp):
    v, t = bumpy_torus(12, 8)
    with pytest.raises(hp.HpsdfError) as e:
        hp.Mesh(v, t[:-1])                       # one triangle missing: an edge has no twin (Mesh.cpp:121-128)
    assert e.value.status == hp.ERR_MESH
    with pytest.raises(hp.HpsdfError):
        hp.Mesh(v, np.array([[0, 1, 10 ** 6]], np.uint32))


def test_octree_of_a_mesh_matches_the_oracle(hp, oracle, torus):
    """C3-style build: mesh SDF -> hp octree, continuity on; same tree, coefficients and queries as the CPU oracle."""
    from oracle import hpref
    v, t, m = torus
    om = oracle.OracleMesh(v, t)
    lo, hi = mesh_root(v)
    kw = dict(threshold=1e-6, nearness=0, strength=0.0, continuity=True, cstrength=8.0, root_min=lo, root_max=hi)
    ocfg = hpref.make_config(threads=8, **kw)
    o = oracle.OracleTree.build(ocfg, hpref.make_program([("mesh", [], om.h)]), threads=8, cg_tol=1e-13)
    cfg = hp.Config(target_error_threshold=1e-6, continuity_enforce=1, continuity_strength=8.0, root_min=lo, root_max=hi)
    tree = hp.Octree()
    tree.Create(cfg, hp.SdfProgram([("mesh", [], m)]), hp.BuildOpts(cg_tolerance=1e-13))
    a, b = hp.parse_block(tree.ToMemoryBlockBytes()), hpref.parse_block(o.block())
    assert a["n_nodes"] == b["n_nodes"] and a["n_coeffs"] == b["n_coeffs"]
    pa, da, ga, ca = leaf_table(a, hp.COEFF_COUNT)
    pb, db, gb, cb = leaf_table(b, hp.COEFF_COUNT)
    ma = {(path_code(p), int(d)): int(g) for p, d, g in zip(pa, da, ga)}
    mb = {(path_code(p), int(d)): int(g) for p, d, g in zip(pb, db, gb)}
    assert not divergent_cells(ma, mb)
    worst = max(rel_inf(x, y) for x, y in zip(ca, cb))
    assert worst <= 1e-10, worst
    pts = np.random.default_rng(3).uniform(lo, hi, (50000, 3))
    assert np.abs(tree.Query(pts) - o.query(pts, 8)).max() <= 1e-9
    st = tree.stats()
    print("mesh octree: nodes", st["n_nodes"], "coeffs", st["n_coeffs"], "fits", st["fits_evaluated"], "sdf evals", st["sdf_evals"],
          "ms", st["total_ms"], "fit ms", st["fit_kernel_ms"], "worst", worst)


@pytest.mark.parametrize("degree", [2, 4, 7])
def test_mixed_mesh_and_closed_form_program_fits_match_the_oracle(hp, oracle, torus, degree):
    """A mesh combined with closed-form primitives goes through the generic sample kernel (interpreter + per-thread BVH
    traversal) instead of meshSampleKernel: min(mesh, sphere) minus a box, fitted on cells near all three surfaces."""
    from oracle import hpref
    v, t, m = torus
    om = oracle.OracleMesh(v, t)
    lo, hi = mesh_root(v)
    items = [("sphere", [0.1, 0.05, 0.0, 0.3]), ("union", []), ("box", [0.3, 0.0, 0.0, 0.1, 0.2, 0.05]), ("subtract", [])]
    prog = hp.SdfProgram([("mesh", [], m)] + items)
    oprog = hpref.make_program([("mesh", [], om.h)] + items)
    cfg = hp.Config(target_error_threshold=1e-6, continuity_enforce=0, root_min=lo, root_max=hi)
    ocfg = hpref.make_config(threshold=1e-6, continuity=False, root_min=lo, root_max=hi, threads=1)
    rng = np.random.default_rng(40 + degree)
    depth = 4
    half = 0.5 ** (depth + 1)
    n = 5 if degree < 7 else 2
    centres = (rng.integers(3, 13, (n, 3)) + 0.5) * 2 * half - 0.5
    cells = np.concatenate([centres, np.full((n, 1), half)], 1).astype(np.float32)
    coeffs, err, _ = hp.fit_batch(cfg, prog, cells, np.full(n, depth), degree)
    for i in range(n):
        c, e = oracle.oracle_fit(ocfg, oprog, centres[i] - half, centres[i] + half, degree, depth)
        assert rel_inf(coeffs[i], c) <= 1e-10
        assert abs(err[i] - e) <= 1e-9 * e + (1e-12 * np.abs(c).max()) ** 2


def test_sample_scratch_chunking_gives_the_same_fits(hp, torus, monkeypatch):
    """Mesh programs sample F into a scratch buffer in chunks of at most 2^26 doubles; force chunks of 3 fits (and of a single
    fit) on a small batch: coefficients and errors must be bitwise those of the one-chunk run, for both sample kernels."""
    v, t, m = torus
    lo, hi = mesh_root(v)
    cfg = hp.Config(target_error_threshold=1e-6, continuity_enforce=0, root_min=lo, root_max=hi)
    rng = np.random.default_rng(9)
    depth, half, n = 4, 0.5 ** 5, 20
    centres = (rng.integers(2, 14, (n, 3)) + 0.5) * 2 * half - 0.5
    cells = np.concatenate([centres, np.full((n, 1), half)], 1).astype(np.float32)
    for items in ([("mesh", [], m)], [("mesh", [], m), ("sphere", [0.1, 0.0, 0.0, 0.25]), ("intersect", [])]):
        prog = hp.SdfProgram(items)
        for degree in (2, 3):
            monkeypatch.delenv("HPSDF_SAMPLE_CAP", raising=False)
            c0, e0, _ = hp.fit_batch(cfg, prog, cells, np.full(n, depth), degree)
            n3 = (4 * degree + 1) ** 3
            for cap in (3 * n3 + 5, 1):
                monkeypatch.setenv("HPSDF_SAMPLE_CAP", str(cap))
                c1, e1, _ = hp.fit_batch(cfg, prog, cells, np.full(n, depth), degree)
                assert np.array_equal(c0, c1) and np.array_equal(e0, e1), (len(items), degree, cap)
    monkeypatch.delenv("HPSDF_SAMPLE_CAP", raising=False)
