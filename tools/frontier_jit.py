"""Frontier fit throughput, interpreted vs run-time specialised kernels (run on the GPU box)."""
import importlib, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
hp = importlib.import_module("hp-adaptive-signed-distance-field-octree_b200")
from common import product_cfg
peak = hp.measure_fp64_peak()
print("fp64 peak TF/s", peak)
for name in ("c2_csg", "sphere_poly_1e8"):
    cfg, prog = product_cfg(hp, name)
    for p in (2, 3, 4, 6):
        row = []
        for jit in (0, 1):
            hp.set_jit(jit)
            fb = hp.bench_frontier(cfg, prog, 5 if p < 6 else 4, p, repeats=3)
            row.append((fb["fits"] / fb["ms_per_launch"] * 1e3, fb["sdf_evals"] / fb["ms_per_launch"] * 1e3, fb["algorithmic_flops"] / fb["ms_per_launch"] * 1e-9 / peak, fb["checksum"]))
        hp.set_jit(0)
        print(name, "p", p, "interp fits/s %.3e evals/s %.3e frac %.3f | jit fits/s %.3e evals/s %.3e frac %.3f | x%.2f checksum rel diff %.2e" %
              (row[0][0], row[0][1], row[0][2], row[1][0], row[1][1], row[1][2], row[1][0] / row[0][0], abs(row[0][3] - row[1][3]) / abs(row[0][3])))
