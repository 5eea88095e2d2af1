import importlib, sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
hp = importlib.import_module("hp-adaptive-signed-distance-field-octree_b200")
from oracle import hporacle, hpref
from cases import *
from common import *
ocfg, oprog = oracle_cfg(hpref, "sphere_poly_1e8")
cfg, prog = product_cfg(hp, "sphere_poly_1e8")
for kw in (dict(max_degree=3), dict(total_mode=1), dict(max_degree=2, max_depth=6)):
  for strict in (0, 1):
    o = hporacle.OracleTree.build(ocfg, oprog, threads=8, **kw)
    t = hp.Octree(); t.Create(cfg, prog, hp.BuildOpts(strict_order=strict, **kw))
    a, b = hp.parse_block(t.ToMemoryBlockBytes()), hpref.parse_block(o.block())
    pa, da, ga, ca = leaf_table(a, hp.COEFF_COUNT); pb, db, gb, cb = leaf_table(b, hp.COEFF_COUNT)
    ma = {(path_code(p), int(d)): int(g) for p, d, g in zip(pa, da, ga)}; mb = {(path_code(p), int(d)): int(g) for p, d, g in zip(pb, db, gb)}
    div = divergent_cells(ma, mb); allowed, nref = logged_cut_group(t)
    st, so = t.stats(), o.stats()
    print(kw, "strict", strict, "nodes", a["n_nodes"], b["n_nodes"], "coeffs", a["n_coeffs"], b["n_coeffs"], "div", len(div), "allowed", len(allowed), "P/H", st["jobs_applied_p"], st["jobs_applied_h"], so["applied_p"], so["applied_h"], "total", st["total_error"], so["final_total"], "rounds", st["rounds"])
    bad = [cell_of(c, d) for c, d in div if cell_of(c, d) not in allowed]
    print("   unlogged:", bad[:6])
    la, lb = t.apply_log(), o.apply_log()
    ea, eb = np.sort(la[:, 3])[::-1], np.sort(lb[:, 3])[::-1]
    n = min(len(ea), len(eb)); d = np.abs(ea[:n] - eb[:n]) / eb[:n]
    print("   sorted initial errs: lens", len(ea), len(eb), "first rel mismatch >1e-9 at", int(np.argmax(d > 1e-9)) if (d > 1e-9).any() else None, "min err applied", ea[-1], eb[-1])
