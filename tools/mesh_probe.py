"""Mesh SDF throughput by query distance (run on the GPU box): python tools/mesh_probe.py U V"""
import importlib, sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
hp = importlib.import_module("hp-adaptive-signed-distance-field-octree_b200")
from meshgen import bumpy_torus, mesh_root
U, V = int(sys.argv[1]), int(sys.argv[2])
verts, tris = bumpy_torus(U, V)
mesh = hp.Mesh(verts, tris)
mn, mx = mesh_root(verts)
rng = np.random.default_rng(0)
n = 1 << 20
uni = rng.uniform(mn, mx, (n, 3)).astype(np.float32)
cent = verts[tris[rng.integers(0, len(tris), n)]].mean(1)
grid = (cent[:n // 512, None, :] + (np.stack(np.meshgrid(*[np.arange(8)] * 3, indexing="ij"), -1).reshape(1, 512, 3) * 1e-3)).reshape(-1, 3).astype(np.float32)
blockrand = (cent[:n // 512, None, :] + rng.uniform(0, 8e-3, (n // 512, 512, 3))).reshape(-1, 3).astype(np.float32)
perm = rng.permutation(len(grid))
for name, pts in [("grid blocks, shuffled", grid[perm]), ("random-in-block, block order", blockrand), ("random-in-block, shuffled", blockrand[perm]),
                  ("grid blocks, first 64k", grid[:65536]), ("grid blocks, 64k shuffled", grid[perm[:65536]]),("uniform in root", uni), ("surface + 1e-3", cent + rng.normal(0, 1e-3, (n, 3)).astype(np.float32)),
                  ("surface + 1e-2", cent + rng.normal(0, 1e-2, (n, 3)).astype(np.float32)),
                  ("surface + 5e-2", cent + rng.normal(0, 5e-2, (n, 3)).astype(np.float32)),
                  ("coherent 8^3 blocks near surface", (cent[:n // 512, None, :] + (np.stack(np.meshgrid(*[np.arange(8)] * 3, indexing="ij"), -1).reshape(1, 512, 3) * 1e-3)).reshape(-1, 3).astype(np.float32))]:
    pts = np.ascontiguousarray(pts, np.float32)
    mesh.SignedDistanceAtPt(pts[:1024])
    t0 = time.perf_counter()
    d = mesh.SignedDistanceAtPt(pts)
    dt = time.perf_counter() - t0
    print("%-36s %.3e evals/s (incl. copies)  mean |d| %.4f" % (name, len(pts) / dt, np.abs(d).mean()), flush=True)
