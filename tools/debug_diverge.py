"""Diff a GPU build against the CPU oracle job by job (run on the GPU box)."""
import importlib, sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
hp = importlib.import_module("hp-adaptive-signed-distance-field-octree_b200")
from oracle import hporacle, hpref
from cases import CASES
from common import product_cfg, oracle_cfg

name = sys.argv[1] if len(sys.argv) > 1 else "csg_small"
cfg, prog = product_cfg(hp, name)
cfg.continuity_enforce = 0
t = hp.Octree(); t.Create(cfg, prog)
ocfg, oprog = oracle_cfg(hpref, name); ocfg.continuity_enforce = 0
o = hporacle.OracleTree.build(ocfg, oprog, threads=8)
a, b = t.apply_log(), o.apply_log()
print("logs", a.shape, b.shape, t.stats())
n = min(len(a), len(b))
same_node = a[:n, 0] == b[:n, 0]
print("first node mismatch at", int(np.argmin(same_node)) if not same_node.all() else None)
rel = np.abs(a[:n, 4] - b[:n, 4]) / np.maximum(np.abs(b[:n, 4]), 1e-300)
print("max rel newErr diff over common prefix (where nodes match):", rel[same_node].max())
i = int(np.argmin(same_node)) if not same_node.all() else n - 1
for k in range(max(0, i - 2), min(n, i + 4)):
    print(k, "gpu", a[k].tolist()); print(k, "cpu", b[k].tolist())
# coarse stage: compare as sets
ca, cb = a[:4096], b[:4096]
oa, ob = np.argsort(ca[:, 0]), np.argsort(cb[:, 0])
print("coarse same node set", np.array_equal(ca[oa, 0], cb[ob, 0]), "max rel err diff", (np.abs(ca[oa, 4] - cb[ob, 4]) / np.maximum(cb[ob, 4], 1e-300)).max(),
      "max abs", np.abs(ca[oa, 4] - cb[ob, 4]).max(), "order equal", np.array_equal(ca[:, 0], cb[:, 0]))
