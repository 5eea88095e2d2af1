import importlib, sys, os, time, ctypes
import numpy as np, torch
ROOT="/root/repo"; sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
hp = importlib.import_module("hp-adaptive-signed-distance-field-octree_b200")
from common import product_cfg
cfg, prog = product_cfg(hp, "c2_csg"); cfg.continuity_enforce = 0
t = hp.Octree(); t.Create(cfg, prog)
for n in (1 << 22, 1 << 24, 100000):
    hpts = torch.empty((n, 3), dtype=torch.float64).pin_memory(); hpts.uniform_(-0.25, 0.5)
    hout = torch.empty(n, dtype=torch.float64).pin_memory()
    for i in range(5):
        if i == 1: torch.cuda.synchronize(); t0 = time.perf_counter()
        hp._check(hp.lib().hpsdf_query(t._h, ctypes.c_void_p(hpts.data_ptr()), n, ctypes.c_void_p(hout.data_ptr())))
    dt = (time.perf_counter() - t0) / 4
    ref = t.Query(hpts[:1000].numpy())
    print(n, "%.3e pts/s" % (n / dt), "ok" if np.array_equal(ref, hout[:1000].numpy()) else "MISMATCH")
