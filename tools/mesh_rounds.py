"""Per-round view of a mesh build (run on the GPU box): python tools/mesh_rounds.py [U V thr maxdeg]   (env HPSDF_MESH_STATS=1 HPSDF_DEBUG_ROUNDS=1)"""
import importlib, sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
hp = importlib.import_module("hp-adaptive-signed-distance-field-octree_b200")
from meshgen import bumpy_torus, mesh_root
U, V = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1000, 435)
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-6
maxdeg = int(sys.argv[4]) if len(sys.argv) > 4 else 11
v, t = bumpy_torus(U, V)
lo, hi = mesh_root(v)
m = hp.Mesh(v, t)
cfg = hp.Config(target_error_threshold=thr, continuity_enforce=0, root_min=lo, root_max=hi)
tree = hp.Octree()
for i in range(3):
    t0 = time.perf_counter()
    tree.Create(cfg, hp.SdfProgram([("mesh", [], m)]), hp.BuildOpts(max_degree=maxdeg, min_round_jobs=int(os.environ.get('MRJ', '0'))))
    s = tree.stats()
    print("Create %.2f ms" % (1e3 * (time.perf_counter() - t0)), {k: (round(s[k], 3) if isinstance(s[k], float) else s[k]) for k in
          ("rounds", "fits_evaluated", "sdf_evals", "fit_kernel_ms", "device_wait_ms", "total_ms", "n_nodes")}, flush=True)
