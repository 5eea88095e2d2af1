"""Helpers shared by the test modules."""
import os

import numpy as np

from cases import CASES, leaf_table, path_code, cell_of, divergent_cells

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


def logged_cut_group(tree):
    """Cells (depth, centre) the build logged as the equal-error group its termination cut falls into, and how many of
    them it refined (north_star: near-threshold refinement divergence is logged with its margin)."""
    cells, refined = set(), 0
    for e in tree.decision_log():
        if e["kind"] in (2, 3):
            cells.add((e["depth"], tuple(float(x) for x in e["centre"])))
            refined += e["kind"] == 2
    return cells, refined


def check_tree_against_golden(block_dict, g, ncount, coeff_tol, tree=None):
    """Compare a parsed MemoryBlock with a golden tree fixture: identical topology, except that leaves inside the
    logged equal-error group at the termination cut (tree.decision_log(), kinds 2/3) may be refined in another order.
    Returns (worst per-leaf coefficient error over the sampled leaves, number of divergent cells)."""
    paths, depth, deg, cs = leaf_table(block_dict, ncount)
    assert block_dict["n_nodes"] == int(g["n_nodes"])
    assert block_dict["n_coeffs"] == int(g["n_coeffs"])
    mine = {(path_code(p), int(d)): int(k) for p, d, k in zip(paths, depth, deg)}
    gold = {(int(c), int(d)): int(k) for c, d, k in zip(g["leaf_code"], g["leaf_depth"], g["leaf_degree"])}
    div = divergent_cells(mine, gold)
    if div:
        assert tree is not None, "topology differs from the golden tree"
        allowed, _ = logged_cut_group(tree)
        for code, d in div:
            assert cell_of(code, d) in allowed, "divergent cell %s is not in the logged near-threshold group" % (cell_of(code, d),)
        # the same number of group members was refined: same degree histogram
        assert np.array_equal(np.bincount(deg, minlength=13), np.bincount(g["leaf_degree"], minlength=13))
    index = {k: i for i, k in enumerate(mine)}
    scale = np.abs(g["leaf_norm"]).max()
    worst, off = 0.0, 0
    gcodes = list(zip(g["leaf_code"], g["leaf_depth"]))
    for li in g["sample_leaves"]:
        key = (int(gcodes[li][0]), int(gcodes[li][1]))
        n = ncount[int(g["leaf_degree"][li])]
        ref = g["sample_coeffs"][off:off + n]
        off += n
        if key in div or key not in index:
            continue
        mine_c = cs[index[key]]
        assert abs(mine_c[0] - g["leaf_c0"][li]) <= coeff_tol * scale
        worst = max(worst, rel_inf(mine_c, ref))
    assert worst <= coeff_tol, "per-leaf |dc|inf/|c|inf = %g" % worst
    return worst, len(div)
