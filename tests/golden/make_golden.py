#!/usr/bin/env python
"""Generate tests/golden/*.npz from the REFERENCE ITSELF compiled here (oracle/_ref/libhpref.so = the unmodified
sources under /root/reference + oracle/eigen_shim + oracle/ref_driver.cpp; parity build, -ffp-contract=off).

Run in the build container (needs /root/reference to have built oracle/_ref):   python tests/golden/make_golden.py
The fixtures travel to the GPU box, where /root/reference does not exist.

Contents per case (deterministic driver: strict greedy, exact-mean nearness, reference totalCoeffError bookkeeping,
CG converged to 1e-13 where continuity is on):
  leaf_depth, leaf_degree, leaf_code   leaves in DFS order by child slot (canonical topology; code = child slots, 3 bits/level)
  leaf_c0, leaf_norm        coeffs[0] and the 2-norm of each leaf's coefficients
  sample_leaves, sample_coeffs   full coefficients of 48 seeded leaves (concatenated)
  query_pts, query_vals     2000 seeded points in (a slightly enlarged) root and the reference's Query values
  n_nodes, n_coeffs, applied_p, applied_h, fits, final_total
sphere_exp_1e8_mc2017.npz: the same contents for the mc_counter nearness mode (seed 2017): 100 calls of the reference's FApprox per fit
on Philox4x32-10 points instead of std::rand() ones.
fits.npz: single reference FitPolynomial calls (coefficients + raw top-shell energy) for seeded cells/degrees.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import hpref            # noqa: E402
from cases import CASES, root_points, leaf_table, path_code  # noqa: E402

TREE_CASES = ["c1_readme", "sphere_poly_1e8", "csg_cont", "custom_domain", "csg_small"]


MC_SEED = 2017          # mc_counter fixture: the reference's 100-sample nearness estimator (through its own FApprox) on Philox points


def tree_golden(name, mc_seed=None, out_name=None):
    c = CASES[name]
    cfg = hpref.make_config(threads=8, **c["cfg"])
    prog = hpref.make_program(c["prog"])
    t = hpref.RefTree.build(cfg, prog, mode=1, threads=8, cg_tol=1e-13, mc_seed=mc_seed)
    blk = hpref.parse_block(t.block())
    paths, depth, deg, cs = leaf_table(blk, hpref.NCOEF)
    rng = np.random.default_rng(1234)
    sample = np.sort(rng.choice(len(cs), size=min(48, len(cs)), replace=False))
    pts = root_points(c["cfg"], 2000, seed=99, margin=0.02)
    st = t.stats()
    np.savez_compressed(os.path.join(HERE, (out_name or name) + ".npz"),
                        leaf_depth=depth.astype(np.uint8), leaf_degree=deg.astype(np.uint8),
                        leaf_code=np.array([path_code(p) for p in paths], np.uint64),
                        leaf_c0=np.array([x[0] for x in cs]), leaf_norm=np.array([np.linalg.norm(x) for x in cs]),
                        sample_leaves=sample, sample_coeffs=np.concatenate([cs[i] for i in sample]),
                        query_pts=pts, query_vals=t.query(pts),
                        n_nodes=blk["n_nodes"], n_coeffs=blk["n_coeffs"], applied_p=st["applied_p"],
                        applied_h=st["applied_h"], fits=st["fits"], final_total=st["final_total"])
    print(out_name or name, blk["n_nodes"], blk["n_coeffs"], st["applied_p"], st["applied_h"])


def fit_golden():
    out = {}
    rng = np.random.default_rng(7)
    k = 0
    for case, degree, depth, degree_in in [("c1_readme", 2, 4, 0), ("c2_csg", 2, 4, 0), ("c2_csg", 3, 5, 0), ("c2_csg", 3, 4, 2),
                                           ("c2_csg", 4, 6, 0), ("sphere_poly_1e8", 4, 4, 3), ("c2_csg", 6, 5, 0),
                                           ("c2_csg", 7, 4, 6), ("sphere_poly_1e8", 5, 7, 0)]:
        c = CASES[case]
        cfg = hpref.make_config(**c["cfg"])
        prog = hpref.make_program(c["prog"])
        for _ in range(6 if degree <= 4 else 2):
            half = 0.5 ** (depth + 1)
            centre = (rng.integers(0, 2 ** depth, 3) + 0.5) * 2 * half - 0.5
            cin = rng.normal(size=hpref.NCOEF[degree_in]) * 1e-3 if degree_in else None
            coeffs, err = hpref.ref_fit(cfg, prog, centre - half, centre + half, degree, depth, degree_in, cin)
            out["fit%03d_meta" % k] = np.array([list(CASES).index(case), degree, depth, degree_in], np.int64)
            out["fit%03d_cell" % k] = np.array([*centre, half], np.float32)
            out["fit%03d_cin" % k] = cin if cin is not None else np.zeros(0)
            out["fit%03d_coeffs" % k] = coeffs
            out["fit%03d_err" % k] = np.array(err)
            k += 1
    out["n_fits"] = np.array(k)
    np.savez_compressed(os.path.join(HERE, "fits.npz"), **out)
    print("fits", k)


def mesh_golden():
    """Reference Mesh::SignedDistanceAtPt (through its own BVH) on the procedural bumpy torus at seeded points."""
    from meshgen import bumpy_torus, mesh_root
    v, t = bumpy_torus(60, 40)
    rm = hpref.RefMesh.create(v, t, bvh=True)
    lo, hi = mesh_root(v)
    pts = np.random.default_rng(42).uniform(lo, hi, (20000, 3)).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "mesh_torus.npz"), pts=pts, sdf=rm.sdf(pts, True, 8))
    print("mesh", len(t))


if __name__ == "__main__":
    assert hpref.available(), "build oracle/_ref first: make -C oracle ref"
    if len(sys.argv) > 1 and sys.argv[1] == "mc":            # only the mc_counter fixture
        tree_golden("sphere_exp_1e8", mc_seed=MC_SEED, out_name="sphere_exp_1e8_mc2017")
        sys.exit(0)
    fit_golden()
    mesh_golden()
    for n in TREE_CASES:
        tree_golden(n)
    tree_golden("sphere_exp_1e8", mc_seed=MC_SEED, out_name="sphere_exp_1e8_mc2017")
