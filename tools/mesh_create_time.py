"""hpsdf_mesh_create timing by stage (run on the GPU box): python tools/mesh_create_time.py [U V]"""
import importlib, sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
hp = importlib.import_module("hp-adaptive-signed-distance-field-octree_b200")
from meshgen import bumpy_torus
U, V = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1000, 435)
v, t = bumpy_torus(U, V)
hp.Mesh(v[:300], np.array([[0, 1, 2], [0, 2, 1]], np.uint32))      # context + first-use costs
for i in range(4):
    if i == 3:
        os.environ["HPSDF_DEBUG_MESH"] = "1"
    t0 = time.perf_counter()
    m = hp.Mesh(v, t)
    print("triangles", len(t), "hpsdf_mesh_create %.2f ms" % (1e3 * (time.perf_counter() - t0)), flush=True)
    m.close()
