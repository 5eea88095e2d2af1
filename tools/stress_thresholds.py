import importlib, sys, os, time
import numpy as np
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
hp = importlib.import_module("hp-adaptive-signed-distance-field-octree_b200")
from cases import CASES
from oracle import hporacle, hpref
k = dict(CASES["sphere_poly_1e8"]["cfg"]); k.setdefault("root_min", (-0.5,) * 3); k.setdefault("root_max", (0.5,) * 3)
for thr in (1e-10, 1e-12, 1e-14):
    cfg = hp.Config(target_error_threshold=thr, nearness_type=0, nearness_strength=0.0, continuity_enforce=0, root_min=k["root_min"], root_max=k["root_max"])
    prog = hp.SdfProgram(CASES["sphere_poly_1e8"]["prog"])
    for jit in (0, 1):
        t = hp.Octree()
        t0 = time.perf_counter(); t.Create(cfg, prog, hp.BuildOpts(jit=jit, total_mode=1)); dt = time.perf_counter() - t0
        s = t.stats()
        pts = np.random.default_rng(0).uniform(-0.5, 0.5, (200000, 3))
        q = t.Query(pts)
        ref = hporacle.sdf_eval(hpref.make_program(CASES["sphere_poly_1e8"]["prog"]), pts)
        print("thr %g jit %d: %.1f ms nodes %d coeffs %d rounds %d fits %d total_err %.3e max|q-F| %.3e rms %.3e" % (thr, jit, dt * 1e3, s["n_nodes"], s["n_coeffs"], s["rounds"], s["fits_evaluated"], s["total_error"], np.abs(q - ref).max(), np.sqrt(np.mean((q - ref) ** 2))), flush=True)
