"""Small end-to-end workload for compute-sanitizer (run on the GPU box):
    compute-sanitizer --tool memcheck python tools/sanitize_target.py"""
import importlib, sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
hp = importlib.import_module("hp-adaptive-signed-distance-field-octree_b200")
from common import product_cfg
from meshgen import bumpy_torus, mesh_root
for name, jit in (("csg_small", 0), ("csg_small", 1), ("csg_cont", 0)):
    cfg, prog = product_cfg(hp, name)
    t = hp.Octree()
    t.Create(cfg, prog, hp.BuildOpts(jit=jit))
    pts = np.random.default_rng(0).uniform(-0.3, 0.6, (100000, 3))
    q = t.Query(pts)
    v, g = t.QueryWithGradient(pts[:1000])
    blk = t.ToMemoryBlockBytes()
    t2 = hp.Octree(); t2.FromMemoryBlock(hp.MemoryBlock.frombytes(blk))
    assert np.array_equal(t2.Query(pts[:5000]), q[:5000])
    print(name, jit, t.stats()["n_nodes"], float(np.nanmin(q)))
v, tr = bumpy_torus(60, 40)
m = hp.Mesh(v, tr)
lo, hi = mesh_root(v)
d = m.SignedDistanceAtPt(np.random.default_rng(1).uniform(lo, hi, (20000, 3)).astype(np.float32))
cfg = hp.Config(target_error_threshold=1e-5, continuity_enforce=1, continuity_strength=8.0, root_min=lo, root_max=hi)
t = hp.Octree(); t.Create(cfg, hp.SdfProgram([("mesh", [], m)]))
t3 = hp.Octree(); t3.Create(cfg, hp.SdfProgram([("mesh", [], m), ("sphere", [0.0, 0.0, 0.0, 0.2]), ("union", [])]))
print("mesh", t.stats()["n_nodes"], t3.stats()["n_nodes"], float(d.min()))
