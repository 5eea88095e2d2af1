"""Workloads shared by the tests, the golden generator and bench.py (BASELINE.json configs)."""
import numpy as np

# name -> (config kwargs in oracle.hpref.make_config vocabulary, SDF program items)
SPHERE_README = [("sphere", [0.0, 0.0, 0.0, 0.25])]                  # README.md:17-20: |p| - 0.25
SPHERE_BENCH = [("sphere", [0.25, 0.0, 0.0, 0.5])]                   # HPUnitTests.cpp:48-51: |p - (0.25,0,0)| - 0.5
# C2: union of box + torus + capsule, f64 closed forms
CSG_C2 = [("box", [0.30, 0.10, 0.10, 0.08, 0.12, 0.06]),
          ("torus", [0.05, 0.20, 0.125, 0.12, 0.04, 1]),
          ("union", []),
          ("capsule", [-0.10, -0.10, 0.05, 0.25, -0.05, 0.30, 0.05]),
          ("union", [])]

CASES = {
    # C1: README config — sphere, thr 1e-6, exponential nearness 3.0, continuity on (lambda 8), root [-0.25,0.5]^3
    "c1_readme": dict(cfg=dict(threshold=1e-6, nearness=2, strength=3.0, continuity=True, cstrength=8.0,
                               root_min=(-0.25,) * 3, root_max=(0.5,) * 3), prog=SPHERE_README),
    # C2: analytic CSG, thr 1e-8, no nearness, continuity off, root [-0.25,0.5]^3
    "c2_csg": dict(cfg=dict(threshold=1e-8, nearness=0, strength=0.0, continuity=False,
                            root_min=(-0.25,) * 3, root_max=(0.5,) * 3), prog=CSG_C2),
    # reference unit test / benchmark sphere (HPUnitTests.cpp:46-77, HPBenchmarks.cpp:25-48)
    "sphere_poly_1e8": dict(cfg=dict(threshold=1e-8, nearness=1, strength=3.0, continuity=False), prog=SPHERE_BENCH),
    "sphere_exp_1e8": dict(cfg=dict(threshold=1e-8, nearness=2, strength=3.0, continuity=False), prog=SPHERE_BENCH),
    "sphere_cont_1e8": dict(cfg=dict(threshold=1e-8, nearness=1, strength=3.0, continuity=True, cstrength=8.0), prog=SPHERE_BENCH),
    # custom domain (HPUnitTests.cpp:285-316): root [-0.25,5]^3, radius 0.75 sphere
    "custom_domain": dict(cfg=dict(threshold=1e-8, nearness=1, strength=3.0, continuity=False,
                                   root_min=(-0.25,) * 3, root_max=(5.0,) * 3), prog=[("sphere", [2.0, 2.0, 2.0, 0.75])]),
    # continuity on a tree with h-splits (mixed-depth faces -> numeric face integrals) and no mirror symmetry
    "csg_cont": dict(cfg=dict(threshold=3e-7, nearness=0, strength=0.0, continuity=True, cstrength=8.0,
                              root_min=(-0.25,) * 3, root_max=(0.5,) * 3), prog=CSG_C2),
    # cheap case with h-splits and several degrees, for fast CPU tests
    "csg_small": dict(cfg=dict(threshold=3e-7, nearness=0, strength=0.0, continuity=False,
                               root_min=(-0.25,) * 3, root_max=(0.5,) * 3), prog=CSG_C2),
}


def root_points(cfg_kwargs, n, seed, margin=0.0):
    """n seeded points uniform in the root box (optionally extended by `margin` of its size to hit the outside)."""
    mn = np.asarray(cfg_kwargs.get("root_min", (-0.5,) * 3), np.float64)
    mx = np.asarray(cfg_kwargs.get("root_max", (0.5,) * 3), np.float64)
    rng = np.random.default_rng(seed)
    ext = (mx - mn) * margin
    return rng.uniform(mn - ext, mx + ext, (n, 3))


def leaf_table(block_dict, ncount):
    """Canonical structural form of a parsed MemoryBlock: leaves in DFS order by child slot ->
    (paths as tuples, depth array, degree array, list of coefficient arrays). Independent of node numbering."""
    nodes, coeffs = block_dict["nodes"], block_dict["coeffs"]
    paths, depths, degs, cs = [], [], [], []
    stack = [(0, ())]
    nochild = np.uint64(0xFFFFFFFFFFFFFFFF)
    while stack:
        idx, path = stack.pop()
        nd = nodes[idx]
        if nd["child"] == nochild:
            d, s = int(nd["deg"]), int(nd["cstart"])
            paths.append(path); depths.append(int(nd["depth"])); degs.append(d); cs.append(coeffs[s:s + ncount[d]])
        else:
            c = int(nd["child"])
            for i in range(7, -1, -1):
                stack.append((c + i, path + (i,)))
    return paths, np.array(depths), np.array(degs), cs


def path_code(path):
    """Child-slot path -> integer (3 bits per level, level 1 in the lowest bits)."""
    code = 0
    for level, c in enumerate(path):
        code |= int(c) << (3 * level)
    return code


def cell_of(code, depth):
    """(depth, centre) of the cell a path code names, in the internal unit cube [-0.5, 0.5]^3."""
    c = np.zeros(3)
    for level in range(depth):
        child = (code >> (3 * level)) & 7
        q = 0.5 ** (level + 2)
        c += np.array([q if child & 1 else -q, q if child & 2 else -q, q if child & 4 else -q])
    return depth, tuple(float(x) for x in c)


def divergent_cells(leaves_a, leaves_b):
    """leaves_*: dict (code, depth) -> degree. Returns the set of (code, depth) cells on which the two trees differ:
    a leaf with another degree, or a leaf in one tree that is split in the other (reported as the coarser cell)."""
    out = set()
    for (la, lb) in ((leaves_a, leaves_b), (leaves_b, leaves_a)):
        for (code, depth), deg in la.items():
            if (code, depth) in lb:
                if lb[(code, depth)] != deg:
                    out.add((code, depth))
                continue
            # not a leaf in the other tree: either an ancestor is a leaf there, or this cell is split there
            anc = [(code & ((1 << (3 * d)) - 1), d) for d in range(depth)]
            hit = [a for a in anc if a in lb]
            out.add(hit[0] if hit else (code, depth))
    return out
