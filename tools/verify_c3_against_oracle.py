"""configs[2] stand-in at full size against the CPU checker (run on the GPU box; the oracle build takes minutes):
870 000-triangle bumpy torus, threshold 1e-6, continuity strength 8 (CG converged to 1e-13 on both sides)."""
import importlib, sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
hp = importlib.import_module("hp-adaptive-signed-distance-field-octree_b200")
from oracle import hporacle, hpref
from cases import leaf_table, path_code, divergent_cells
from common import rel_inf, logged_cut_group
from meshgen import bumpy_torus, mesh_root
U, V = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1000, 435)
THR = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-6
CONT = int(sys.argv[4]) if len(sys.argv) > 4 else 1
MAXDEG = int(sys.argv[5]) if len(sys.argv) > 5 else 11
v, t = bumpy_torus(U, V)
lo, hi = mesh_root(v)
m = hp.Mesh(v, t)
om = hporacle.OracleMesh(v, t)
threads = os.cpu_count() or 1
kw = dict(threshold=THR, nearness=0, strength=0.0, continuity=bool(CONT), cstrength=8.0, root_min=lo, root_max=hi)
t0 = time.perf_counter()
o = hporacle.OracleTree.build(hpref.make_config(threads=threads, **kw), hpref.make_program([("mesh", [], om.h)]), threads=threads, cg_tol=1e-13, max_degree=MAXDEG)
t_oracle = time.perf_counter() - t0
cfg = hp.Config(target_error_threshold=THR, continuity_enforce=CONT, continuity_strength=8.0, root_min=lo, root_max=hi)
tree = hp.Octree()
tree.Create(cfg, hp.SdfProgram([("mesh", [], m)]), hp.BuildOpts(cg_tolerance=1e-13, max_degree=MAXDEG))
t0 = time.perf_counter()
tree.Create(cfg, hp.SdfProgram([("mesh", [], m)]), hp.BuildOpts(cg_tolerance=1e-13, max_degree=MAXDEG))
t_gpu = time.perf_counter() - t0
a, b = hp.parse_block(tree.ToMemoryBlockBytes()), hpref.parse_block(o.block())
pa, da, ga, ca = leaf_table(a, hp.COEFF_COUNT)
pb, db, gb, cb = leaf_table(b, hp.COEFF_COUNT)
ma = {(path_code(p), int(d)): i for i, (p, d) in enumerate(zip(pa, da))}
mb = {(path_code(p), int(d)): i for i, (p, d) in enumerate(zip(pb, db))}
div = divergent_cells({k: int(ga[i]) for k, i in ma.items()}, {k: int(gb[i]) for k, i in mb.items()})
worst = max(rel_inf(ca[i], cb[mb[k]]) for k, i in ma.items() if k in mb and k not in div)
pts = np.random.default_rng(3).uniform(lo, hi, (200000, 3))
dq = np.abs(tree.Query(pts) - o.query(pts, threads)).max()
print("threshold %g continuity %d max degree %d | triangles %d | oracle Create %.1f s on %d threads, GPU Create %.1f ms | nodes %d / %d coeffs %d / %d | divergent cells %d (logged tie group: %d) | "
      "worst per-leaf |dc|inf/|c|inf %.2e | max |dQuery| over 2e5 points %.2e" %
      (THR, CONT, MAXDEG, len(t), t_oracle, threads, t_gpu * 1e3, a["n_nodes"], b["n_nodes"], a["n_coeffs"], b["n_coeffs"], len(div), len(logged_cut_group(tree)[0]), worst, dq))
