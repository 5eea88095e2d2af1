"""Query kernel micro-benchmark (GPU box): python tools/bench_query.py [case]"""
import importlib, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
hp = importlib.import_module("hp-adaptive-signed-distance-field-octree_b200")
from common import product_cfg
from cases import CASES
name = sys.argv[1] if len(sys.argv) > 1 else "c2_csg"
cfg, prog = product_cfg(hp, name); cfg.continuity_enforce = 0
t = hp.Octree(); t.Create(cfg, prog)
n = 1 << 24
mn, mx = CASES[name]["cfg"].get("root_min", (-0.5,)*3)[0], CASES[name]["cfg"].get("root_max", (0.5,)*3)[0]
pts = (torch.rand((n, 3), device="cuda", dtype=torch.float64) * (mx - mn) + mn).contiguous()
out = torch.empty(n, device="cuda", dtype=torch.float64)
s = torch.cuda.current_stream().cuda_stream
for _ in range(3): t.QueryDevice(pts.data_ptr(), n, out.data_ptr(), s)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): t.QueryDevice(pts.data_ptr(), n, out.data_ptr(), s)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(os.environ.get("HPSDF_LIB", "default"), name, "%.3f ms  %.3e pts/s  checksum %.6f" % (ms, n / ms * 1e3, float(out.sum())))
