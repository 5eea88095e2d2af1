"""C2 Create time against min_round_jobs (run on the GPU box)."""
import importlib, sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
hp = importlib.import_module("hp-adaptive-signed-distance-field-octree_b200")
from common import product_cfg
cfg, prog = product_cfg(hp, sys.argv[1] if len(sys.argv) > 1 else "c2_csg")
t = hp.Octree()
for mrj in (256, 512, 1024, 2048, 4096, 8192):
    best = 1e9
    for i in range(6):
        t0 = time.perf_counter()
        t.Create(cfg, prog, hp.BuildOpts(jit=1, min_round_jobs=mrj))
        best = min(best, 1e3 * (time.perf_counter() - t0))
    s = t.stats()
    print("min_round_jobs", mrj, "best wall %.3f ms" % best, {k: (round(s[k], 3) if isinstance(s[k], float) else s[k]) for k in ("rounds", "fits_evaluated", "fit_kernel_ms", "device_wait_ms", "pack_ms", "finalize_ms", "n_nodes")}, flush=True)
