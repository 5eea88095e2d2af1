#!/usr/bin/env python
"""Print every build statistic of three consecutive C4 Creates (where the time of one Create goes)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
cx = bench.Ctx()
hp, torch = cx.hp, cx.torch
verts, tris, box, cfg = bench.mesh_case(cx, bench.MESH_C4_UV, bench.C4)
mesh = hp.Mesh(verts, tris, device=cx.local)
prog = hp.SdfProgram([("mesh", [], mesh)])
tree = hp.Octree()
for i in range(3):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record()
    tree.Create(cfg, prog, cx.opts(max_degree=bench.C4["max_degree"]))
    t1 = time.perf_counter(); b.record(); torch.cuda.synchronize()
    print("create %d: wall %.1f ms, events %.1f ms" % (i, 1e3 * (t1 - t0), a.elapsed_time(b)))
    print(json.dumps({k: (round(v, 3) if isinstance(v, float) else v) for k, v in tree.stats().items()}))
