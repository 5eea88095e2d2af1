"""Print the phase timers of warm builds (run on the GPU box): python tools/profile_build.py [case] [speculate] [jit]"""
import importlib, sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
hp = importlib.import_module("hp-adaptive-signed-distance-field-octree_b200")
from common import product_cfg
name = sys.argv[1] if len(sys.argv) > 1 else "c2_csg"
spec = int(sys.argv[2]) if len(sys.argv) > 2 else 0
jit = int(sys.argv[3]) if len(sys.argv) > 3 else 0
cfg, prog = product_cfg(hp, name)
t = hp.Octree()
keys = ["rounds", "fits_evaluated", "jobs_evaluated", "kernel_launches", "total_ms", "fit_kernel_ms", "device_wait_ms", "host_replay_ms",
        "host_select_ms", "host_tasks_ms", "pack_ms", "finalize_ms", "continuity_ms", "continuity_enum_ms", "continuity_assembly_ms", "continuity_cg_ms", "cg_iterations", "n_nodes", "n_coeffs"]
for i in range(4):
    t0 = time.perf_counter()
    t.Create(cfg, prog, hp.BuildOpts(speculate=spec, jit=jit))
    wall = 1e3 * (time.perf_counter() - t0)
    s = t.stats()
    print("run", i, "wall %.2f ms" % wall, {k: (round(s[k], 3) if isinstance(s[k], float) else s[k]) for k in keys})
