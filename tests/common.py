"""Helpers shared by the test modules."""
import os

import numpy as np

from cases import CASES, leaf_table

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NEAR = dict(none=0, poly=1, exp=2)


def oracle_cfg(hpref_mod, name, threads=8):
    c = CASES[name]
    return hpref_mod.make_config(threads=threads, **c["cfg"]), hpref_mod.make_program(c["prog"])


def product_cfg(hp, name):
    """hp.Config / hp.SdfProgram of a named case."""
    c = CASES[name]
    k = c["cfg"]
    cfg = hp.Config(target_error_threshold=k["threshold"], nearness_type=k.get("nearness", 0),
                    nearness_strength=k.get("strength", 0.0), continuity_enforce=1 if k.get("continuity", True) else 0,
                    continuity_strength=k.get("cstrength", 8.0), thread_count=8,
                    root_min=k.get("root_min", (-0.5,) * 3), root_max=k.get("root_max", (0.5,) * 3))
    return cfg, hp.SdfProgram(c["prog"])


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def rel_inf(a, b):
    """Per-node coefficient error: |a-b|_inf / |b|_inf (element-wise relative error is meaningless for
    high-order coefficients that are 1e-6..1e-10 of c000)."""
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def check_tree_against_golden(block_dict, g, ncount, coeff_tol, topo_exact=True):
    """Compare a parsed MemoryBlock with a golden tree fixture. Returns the worst per-leaf coefficient error."""
    paths, depth, deg, cs = leaf_table(block_dict, ncount)
    assert block_dict["n_nodes"] == int(g["n_nodes"])
    assert block_dict["n_coeffs"] == int(g["n_coeffs"])
    assert np.array_equal(depth, g["leaf_depth"]), "leaf depths differ (topology)"
    assert np.array_equal(deg, g["leaf_degree"]), "leaf degrees differ (topology)"
    c0 = np.array([x[0] for x in cs])
    scale = np.abs(g["leaf_norm"]).max()
    assert np.abs(c0 - g["leaf_c0"]).max() <= coeff_tol * scale
    worst, off = 0.0, 0
    for i in g["sample_leaves"]:
        n = ncount[deg[i]]
        worst = max(worst, rel_inf(cs[i], g["sample_coeffs"][off:off + n]))
        off += n
    assert worst <= coeff_tol, "per-leaf |dc|inf/|c|inf = %g" % worst
    return worst
