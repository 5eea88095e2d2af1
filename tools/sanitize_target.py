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
# round 2: small mesh launches (the tail phase of meshSampleKernel dominates them), mc_counter nearness, the host scheduler,
# ray queries, the boolean operations, the device point generator
c0 = 0.5 * (np.asarray(lo) + np.asarray(hi)); ext = float(np.max(np.asarray(hi) - np.asarray(lo)))
cells = np.array([[0.0, 0.0, 0.0, 0.125], [0.25, 0.0, 0.0, 0.0625]] * 8, np.float32)      # internal unit-cube cells
for degree in (2, 4):
    hp.fit_batch(cfg, hp.SdfProgram([("mesh", [], m)]), cells, np.full(len(cells), 2), degree)
cfg2, prog2 = product_cfg(hp, "sphere_exp_1e8")
cfg2.target_error_threshold = 1e-6
t4 = hp.Octree(); t4.Create(cfg2, prog2, hp.BuildOpts(nearness_mode=hp.NEARNESS_MC_COUNTER, nearness_seed=7))
t5 = hp.Octree(); t5.Create(cfg2, prog2, hp.BuildOpts(scheduler=1))
t4.UnionSDF(hp.SdfProgram([("sphere", [0.2, 0.0, 0.0, 0.2])]))
o = np.random.default_rng(2).uniform(-0.4, 0.4, (2000, 3)); dirs = np.random.default_rng(3).normal(size=(2000, 3))
hit = t4.QueryRay(o, dirs / np.linalg.norm(dirs, axis=1, keepdims=True), 2.0)
import torch
buf = torch.empty((4096, 3), dtype=torch.float64, device="cuda")
hp.uniform_points_device(123, 0, 4096, (-0.5,) * 3, (0.5,) * 3, buf.data_ptr())
torch.cuda.synchronize()
print("round 2:", t4.stats()["n_nodes"], t5.stats()["n_nodes"], int(np.asarray(hit[0]).sum()), float(buf.sum()))
