"""The mesh configs at scale (run on the GPU box): python tools/mesh_scale.py U V threshold [continuity] [query_points]
U=1000 V=435 is the 870 k-triangle dragon stand-in (configs[2]), U=1000 V=800 the 1.6 M-triangle Ramesses stand-in."""
import importlib, sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
hp = importlib.import_module("hp-adaptive-signed-distance-field-octree_b200")
from meshgen import bumpy_torus, mesh_root
U, V = int(sys.argv[1]), int(sys.argv[2])
thr = float(sys.argv[3])
cont = int(sys.argv[4]) if len(sys.argv) > 4 else 0
nq = int(float(sys.argv[5])) if len(sys.argv) > 5 else 0
maxdeg = int(sys.argv[6]) if len(sys.argv) > 6 else 11
minjobs = int(sys.argv[7]) if len(sys.argv) > 7 else 0
verts, tris = bumpy_torus(U, V)
t0 = time.perf_counter()
mesh = hp.Mesh(verts, tris)
print("mesh: %d triangles, create %.1f ms" % (len(tris), 1e3 * (time.perf_counter() - t0)), flush=True)
mn, mx = mesh_root(verts)
cfg = hp.Config(target_error_threshold=thr, nearness_type=0, nearness_strength=0.0, continuity_enforce=cont, continuity_strength=8.0,
                thread_count=8, root_min=mn, root_max=mx)
prog = hp.SdfProgram([("mesh", [], mesh)])
keys = ["rounds", "fits_evaluated", "sdf_evals", "total_ms", "fit_kernel_ms", "device_wait_ms", "host_replay_ms", "host_select_ms", "host_tasks_ms",
        "pack_ms", "continuity_ms", "continuity_cg_ms", "cg_iterations", "n_nodes", "n_coeffs"]
for i in range(2):
    t = hp.Octree()
    t0 = time.perf_counter()
    t.Create(cfg, prog, hp.BuildOpts(max_degree=maxdeg, min_round_jobs=minjobs))
    wall = 1e3 * (time.perf_counter() - t0)
    s = t.stats()
    print("run", i, "wall %.1f ms" % wall, {k: (round(s[k], 3) if isinstance(s[k], float) else s[k]) for k in keys}, flush=True)
print("mesh sdf evals/s inside Create: %.3e" % (s["sdf_evals"] / s["fit_kernel_ms"] * 1e3))
if nq:
    import torch
    pts = (torch.rand((nq, 3), device="cuda", dtype=torch.float64) * (mx[0] - mn[0]) + torch.tensor(mn, device="cuda", dtype=torch.float64)).contiguous()
    out = torch.empty(nq, device="cuda", dtype=torch.float64)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        t.QueryDevice(pts.data_ptr(), nq, out.data_ptr(), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        t.QueryDevice(pts.data_ptr(), nq, out.data_ptr(), st)
    e1.record(); torch.cuda.synchronize()
    print("query %.3e points/s" % (5 * nq / (e0.elapsed_time(e1) * 1e-3)))
