#!/usr/bin/env python
"""Duration of SMALL mesh fit launches (what one rank of eight sees in a late round of the 870 k-triangle build): the deepest
leaves of the built tree are refitted in batches of 8..2048 cells, with and without the tail-phase hand-over."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
cx = bench.Ctx()
hp = cx.hp
verts, tris, box, cfg = bench.mesh_case(cx, bench.MESH_UV, bench.C3)
mesh = hp.Mesh(verts, tris, device=cx.local)
prog = hp.SdfProgram([("mesh", [], mesh)])
tree = hp.Octree()
tree.Create(cfg, prog, cx.opts())
blk = hp.parse_block(tree.ToMemoryBlockBytes())
nd = blk["nodes"]
leaf = nd["deg"] != 13
depth = nd["depth"][leaf]
mn, mx = nd["mn"][leaf], nd["mx"][leaf]
order = np.argsort(-depth.astype(int), kind="stable")
cells = np.concatenate([(mn + mx) * 0.5, ((mx - mn) * 0.5)[:, :1]], axis=1).astype(np.float32)[order]
depth = depth[order]
print("leaves", leaf.sum(), "deepest", depth[:4], "shallowest used", depth[2047])
for degree in ((2,) if os.environ.get("HPSDF_MESH_STATS") else (2, 3, 4)):
    for n in ((8, 128, 512) if os.environ.get("HPSDF_MESH_STATS") else (8, 32, 128, 512, 2048)):
        row = []
        for off in ("", "1"):
            if off:
                os.environ["HPSDF_MESH_NO_HANDOVER"] = "1"
            else:
                os.environ.pop("HPSDF_MESH_NO_HANDOVER", None)
            best = 1e9
            ref = None
            for rep in range(1 if os.environ.get("HPSDF_MESH_STATS") else 4):
                c, e, ms = hp.fit_batch(cfg, prog, cells[:n], depth[:n], degree)
                best = min(best, ms)
            row.append((best, c))
        same = np.array_equal(row[0][1], row[1][1])
        print("degree %d, %4d fits, %8d samples: hand-over %.3f ms, without %.3f ms, identical coefficients: %s" % (degree, n, n * (4 * degree + 1) ** 3, row[0][0], row[1][0], same))
os.environ.pop("HPSDF_MESH_NO_HANDOVER", None)
