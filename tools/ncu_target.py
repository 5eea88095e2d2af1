"""Short single-GPU workloads for `ncu --set full` captures: python tools/ncu_target.py fit|query|continuity"""
import importlib, sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
hp = importlib.import_module("hp-adaptive-signed-distance-field-octree_b200")
from common import product_cfg
what = sys.argv[1] if len(sys.argv) > 1 else "fit"
cfg, prog = product_cfg(hp, "c2_csg")
if what == "fitjit":
    hp.set_jit(True)
    for p in (2, 3):
        print(p, hp.bench_frontier(cfg, prog, 5, p, repeats=1))
elif what == "fit":
    for p in (2, 3):
        print(p, hp.bench_frontier(cfg, prog, 5, p, repeats=1))
elif what == "mesh":
    from meshgen import bumpy_torus, mesh_root
    v, t = bumpy_torus(1000, 435)
    lo, hi = mesh_root(v)
    m = hp.Mesh(v, t)
    mcfg = hp.Config(target_error_threshold=1e-6, continuity_enforce=1, continuity_strength=8.0, root_min=lo, root_max=hi)
    tree = hp.Octree()
    tree.Create(mcfg, hp.SdfProgram([("mesh", [], m)]))
    print(tree.stats())
elif what == "sched":
    t = hp.Octree()
    for _ in range(2):
        t.Create(cfg, prog, hp.BuildOpts(jit=1))
    print(t.stats())
elif what == "query":
    import torch
    t = hp.Octree(); t.Create(cfg, prog)
    n = 1 << 24
    pts = (torch.rand((n, 3), device="cuda", dtype=torch.float64) * 0.75 - 0.25).contiguous()
    out = torch.empty(n, device="cuda", dtype=torch.float64)
    for _ in range(3):
        t.QueryDevice(pts.data_ptr(), n, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    print(float(out.sum()))
else:
    cfg, prog = product_cfg(hp, "csg_cont")
    t = hp.Octree(); t.Create(cfg, prog)
    print(t.stats())
