"""Device-resident scheduler against the host replay (run on the GPU box): python tools/sched_check.py [case ...]
For each case: canonical tree + coefficients of both schedulers must agree (outside the logged tie group), timings printed."""
import importlib, sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
hp = importlib.import_module("hp-adaptive-signed-distance-field-octree_b200")
from cases import CASES, leaf_table, path_code, cell_of, divergent_cells
from common import product_cfg, logged_cut_group, rel_inf
names = sys.argv[1:] or ["csg_small", "sphere_poly_1e8", "c2_csg", "sphere_exp_1e8", "custom_domain", "c1_readme"]
keys = ["rounds", "fits_evaluated", "jobs_evaluated", "jobs_applied_p", "jobs_applied_h", "kernel_launches", "total_ms", "fit_kernel_ms", "device_wait_ms",
        "host_replay_ms", "pack_ms", "finalize_ms", "n_nodes", "n_coeffs", "total_error", "cut_margin"]
for name in names:
    cfg, prog = product_cfg(hp, name)
    trees = {}
    for sched in (1, 0):
        t = hp.Octree()
        for i in range(3):
            t0 = time.perf_counter()
            t.Create(cfg, prog, hp.BuildOpts(scheduler=sched, jit=1))
            wall = 1e3 * (time.perf_counter() - t0)
        s = t.stats()
        print(name, "scheduler", "host" if sched else "device", "wall %.3f ms" % wall, {k: (round(s[k], 4) if isinstance(s[k], float) else s[k]) for k in keys}, flush=True)
        trees[sched] = t
    a, b = hp.parse_block(trees[0].ToMemoryBlockBytes()), hp.parse_block(trees[1].ToMemoryBlockBytes())
    ok = a["n_nodes"] == b["n_nodes"] and a["n_coeffs"] == b["n_coeffs"]
    if ok:
        pa, da, ga, ca = leaf_table(a, hp.COEFF_COUNT)
        pb, db, gb, cb = leaf_table(b, hp.COEFF_COUNT)
        ma = {(path_code(p), int(d)): i for i, (p, d) in enumerate(zip(pa, da))}
        mb = {(path_code(p), int(d)): i for i, (p, d) in enumerate(zip(pb, db))}
        div = divergent_cells({k: int(ga[i]) for k, i in ma.items()}, {k: int(gb[i]) for k, i in mb.items()})
        allowed = logged_cut_group(trees[0])[0] | logged_cut_group(trees[1])[0]
        unlogged = [cell_of(c, d) for c, d in div if cell_of(c, d) not in allowed]
        worst = max(rel_inf(ca[i], cb[mb[k]]) for k, i in ma.items() if k in mb and k not in div)
        print("  ->", name, "nodes", a["n_nodes"], "divergent", len(div), "unlogged", len(unlogged), "worst coeff diff", worst, "PASS" if not unlogged and worst <= 1e-12 else "FAIL", flush=True)
    else:
        print("  ->", name, "FAIL: nodes", a["n_nodes"], b["n_nodes"], "coeffs", a["n_coeffs"], b["n_coeffs"], flush=True)
