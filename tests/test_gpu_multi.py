"""GPU, 2 ranks over NCCL (skipped on a single-GPU box; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`).

One tree built cooperatively by two ranks must be the tree one rank builds alone: the sharded rounds (mesh program: every
round; closed-form program: forced by min shard size through a tiny threshold is not needed — the mesh case covers the
exchange) replicate coefficients and records with grouped broadcasts, every rank replays the same greedy order. Jobs are
dealt out by a stride permutation at world > 1, so node NUMBERING differs; the canonical tree (paths, degrees, coefficients)
must be identical, bit for bit in the coefficients."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import importlib
    import torch
    import torch.distributed as dist
    from cases import CASES, leaf_table, path_code
    from common import product_cfg
    from meshgen import bumpy_torus, mesh_root
    hp = importlib.import_module("hp-adaptive-signed-distance-field-octree_b200")
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(hp.Comm.unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    comm = hp.Comm(bytes(uid.cpu().numpy().tobytes()), rank, world, rank)

    def canon(tree):
        blk = hp.parse_block(tree.ToMemoryBlockBytes())
        paths, depth, deg, cs = leaf_table(blk, hp.COEFF_COUNT)
        return {(path_code(p), int(d)): (int(g), np.asarray(c)) for p, d, g, c in zip(paths, depth, deg, cs)}, blk["n_nodes"], blk["n_coeffs"]

    out = {}
    v, t = bumpy_torus(60, 40)
    mesh = hp.Mesh(v, t, device=rank)
    lo, hi = mesh_root(v)
    cases = {"mesh": (hp.Config(target_error_threshold=1e-6, continuity_enforce=1, continuity_strength=8.0, root_min=lo, root_max=hi),
                      hp.SdfProgram([("mesh", [], mesh)])),
             "csg_small": product_cfg(hp, "csg_small")}
    for name, (cfg, prog) in cases.items():
        solo = hp.Octree()
        solo.Create(cfg, prog, hp.BuildOpts(device=rank))
        both = hp.Octree()
        o = hp.BuildOpts(device=rank)
        o.comm = comm._h
        both.Create(cfg, prog, o)
        a, na, ca = canon(solo)
        b, nb, cb = canon(both)
        same = na == nb and ca == cb and a.keys() == b.keys() and all(a[k][0] == b[k][0] and np.array_equal(a[k][1], b[k][1]) for k in a)
        pts = np.random.default_rng(2).uniform(lo, hi, (20000, 3)) if name == "mesh" else np.random.default_rng(2).uniform(-0.25, 0.5, (20000, 3))
        out[name] = bool(same) and bool(np.array_equal(solo.Query(pts), both.Query(pts)))
    ret[rank] = out
    comm.close()
    dist.destroy_process_group()


def test_two_rank_build_equals_single_rank_build():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29600 + os.getpid() % 200
    mp.spawn(worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] == {"mesh": True, "csg_small": True}, dict(ret)
    assert ret[1] == ret[0]
