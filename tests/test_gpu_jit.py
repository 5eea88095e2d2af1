"""Run-time specialised fit kernels (hpsdf_build_opts.jit, csrc/jit.cpp) against the same checkers as the interpreted
kernels: the golden trees of the reference, the oracle's fits, and the interpreted kernels themselves."""
import numpy as np
import pytest

from cases import CASES
from common import check_tree_against_golden, golden, oracle_cfg, product_cfg, rel_inf

COEFF_TOL = 1e-10
QUERY_TOL = 1e-9

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("degree", [1, 2, 3, 6, 7, 11])
def test_jit_fit_matches_oracle_and_interpreted(hp, oracle, degree):
    from oracle import hpref
    cfg, prog = product_cfg(hp, "c2_csg")
    ocfg, oprog = oracle_cfg(hpref, "c2_csg")
    rng = np.random.default_rng(100 + degree)
    depth = 5
    half = 0.5 ** (depth + 1)
    n = 6 if degree <= 6 else 2
    centres = (rng.integers(8, 24, (n, 3)) + 0.5) * 2 * half - 0.5
    cells = np.concatenate([centres, np.full((n, 1), half)], 1).astype(np.float32)
    ref_c, ref_e, _ = hp.fit_batch(cfg, prog, cells, np.full(n, depth), degree)
    hp.set_jit(True)
    try:
        coeffs, err, _ = hp.fit_batch(cfg, prog, cells, np.full(n, depth), degree)
    finally:
        hp.set_jit(False)
    for i in range(n):
        assert rel_inf(coeffs[i], ref_c[i]) <= 1e-13          # same arithmetic up to FMA contraction of a few constants
        c, e = oracle.oracle_fit(ocfg, oprog, centres[i] - half, centres[i] + half, degree, depth)
        assert rel_inf(coeffs[i], c) <= COEFF_TOL
        assert abs(err[i] - e) <= 1e-9 * e + (1e-12 * np.abs(c).max()) ** 2


@pytest.mark.parametrize("name", ["sphere_poly_1e8", "csg_small", "custom_domain"])
def test_jit_create_matches_reference_golden(hp, name):
    cfg, prog = product_cfg(hp, name)
    t = hp.Octree()
    t.Create(cfg, prog, hp.BuildOpts(jit=1))
    g = golden(name)
    blk = hp.parse_block(t.ToMemoryBlockBytes())
    worst, ndiv = check_tree_against_golden(blk, g, hp.COEFF_COUNT, COEFF_TOL, tree=t)
    st = t.stats()
    assert st["jobs_applied_p"] == int(g["applied_p"]) and st["jobs_applied_h"] == int(g["applied_h"])
    if ndiv == 0:
        assert np.abs(t.Query(g["query_pts"]) - g["query_vals"]).max() <= QUERY_TOL
    print(name, "jit: worst", worst, "divergent", ndiv, "ms", st["total_ms"])


def test_jit_all_primitives_match_interpreted(hp):
    """Every closed-form opcode through the generator (sphere, box, 3 torus axes, capsule, plane, all operators)."""
    items = [("box", [0.1, 0.0, 0.0, 0.2, 0.3, 0.1]), ("sphere", [0.2, 0.1, 0, 0.25]), ("subtract", []),
             ("plane", [0.0, 0.6, 0.8, -0.05]), ("intersect", []), ("torus", [0, 0, 0, 0.3, 0.05, 2]), ("union", []),
             ("torus", [0.1, 0, 0, 0.2, 0.04, 0]), ("union", []), ("torus", [0, 0.1, 0, 0.25, 0.03, 1]), ("union", []),
             ("capsule", [-0.2, -0.1, 0.05, 0.25, -0.05, 0.3, 0.05]), ("union", []), ("negate", [])]
    cfg = hp.Config()
    prog = hp.SdfProgram(items)
    rng = np.random.default_rng(7)
    depth, half = 4, 0.5 ** 5
    centres = (rng.integers(0, 16, (64, 3)) + 0.5) * 2 * half - 0.5
    cells = np.concatenate([centres, np.full((64, 1), half)], 1).astype(np.float32)
    a, ea, _ = hp.fit_batch(cfg, prog, cells, np.full(64, depth), 3)
    hp.set_jit(True)
    try:
        b, eb, _ = hp.fit_batch(cfg, prog, cells, np.full(64, depth), 3)
    finally:
        hp.set_jit(False)
    assert max(rel_inf(b[i], a[i]) for i in range(64)) <= 1e-13


def test_jit_refuses_nothing_silently(hp):
    """Mesh / octree programs are documented to stay interpreted even with jit=1; closed-form ones must really be JIT-built."""
    src, nbytes = hp.jit_compile_check(hp.SdfProgram(CASES["c2_csg"]["prog"]), 4)
    assert "fit_kernel_body.cuh" in src and nbytes > 10000
