import importlib, sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
hp = importlib.import_module("hp-adaptive-signed-distance-field-octree_b200")
from oracle import hporacle, hpref
from common import *
name = sys.argv[1] if len(sys.argv) > 1 else "csg_small"
strict = int(sys.argv[2]) if len(sys.argv) > 2 else 0
ocfg, oprog = oracle_cfg(hpref, name); ocfg.continuity_enforce = 0
cfg, prog = product_cfg(hp, name); cfg.continuity_enforce = 0
o = hporacle.OracleTree.build(ocfg, oprog, threads=8)
t = hp.Octree(); t.Create(cfg, prog, hp.BuildOpts(strict_order=strict))
la, lb = t.apply_log(), o.apply_log()
print("jobs", len(la), len(lb), t.stats()["total_error"], o.stats()["final_total"], "rounds", t.stats()["rounds"])
print("gpu last 3:", la[-3:, [0,1,2,3,4,7]].tolist())
print("cpu last 3:", lb[-3:, [0,1,2,3,4,7]].tolist())
ea, eb = np.sort(la[:, 3])[::-1], np.sort(lb[:, 3])[::-1]
n = min(len(ea), len(eb)); d = np.abs(ea[:n] - eb[:n]) / eb[:n]
i = int(np.argmax(d > 1e-9)) if (d > 1e-9).any() else None
print("first sorted mismatch", i, None if i is None else (ea[i-2:i+3].tolist(), eb[i-2:i+3].tolist()))
print("min applied err gpu/cpu", ea[-1], eb[-1])
# is the gpu log monotone in strict mode? number of inversions of initial_err after coarse
ia = la[4096:, 3]
print("gpu non-coarse: count where next err > prev err:", int((np.diff(ia) > 0).sum()), "of", len(ia))
ib = lb[4096:, 3]
print("cpu non-coarse: count where next err > prev err:", int((np.diff(ib) > 0).sum()), "of", len(ib))
# total trajectory
print("gpu total at job 4096:", la[4095, 7], "cpu:", lb[4095, 7])
