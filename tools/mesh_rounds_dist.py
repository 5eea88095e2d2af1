"""Per-round view of a sharded mesh build: torchrun --nproc-per-node N tools/mesh_rounds_dist.py   (env HPSDF_DEBUG_ROUNDS=1 on rank 0)"""
import importlib, sys, os, time
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
hp = importlib.import_module("hp-adaptive-signed-distance-field-octree_b200")
from meshgen import bumpy_torus, mesh_root
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    uid.copy_(torch.frombuffer(bytearray(hp.Comm.unique_id()), dtype=torch.uint8))
dist.broadcast(uid, 0)
comm = hp.Comm(bytes(uid.cpu().numpy().tobytes()), rank, world, local)
v, t = bumpy_torus(1000, 435)
lo, hi = mesh_root(v)
m = hp.Mesh(v, t, device=local)
cfg = hp.Config(target_error_threshold=1e-6, continuity_enforce=0, root_min=lo, root_max=hi)
opts = hp.BuildOpts(device=local, min_round_jobs=int(os.environ.get('MRJ', '0')))
opts.comm = comm._h
tree = hp.Octree()
for i in range(3):
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    tree.Create(cfg, hp.SdfProgram([("mesh", [], m)]), opts)
    if True:
        s = tree.stats()
        print("rank", rank, "Create %.2f ms" % (1e3 * (time.perf_counter() - t0)), {k: (round(s[k], 3) if isinstance(s[k], float) else s[k]) for k in ("rounds", "fit_kernel_ms", "device_wait_ms", "total_ms")}, flush=True)
comm.close(); dist.destroy_process_group()
