"""CPU, world_size 2 over gloo: the host-side convention of the sharded build.

In a multi-GPU Create every rank builds the same task list per round, evaluates the contiguous shard
hpsdf_shard_range(n, rank, world) of every degree group and receives the other shards (one grouped NCCL broadcast per
round on the GPU; an all_gather here), so that all ranks hold identical records and replay the same greedy order.
This test runs that convention with the CPU oracle standing in for the fit kernel: each rank fits only its shard of the
4096 coarse cells of C2, the shards are exchanged, and the merged records must equal a single-process evaluation
bit for bit, on every rank."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def coarse_cells(n):
    """First n depth-4 cells in UniformlyRefine order (any fixed order works for this test)."""
    cells = []
    h = 0.5 ** 5
    for k in range(n):
        ix, iy, iz = k % 16, (k // 16) % 16, (k // 256) % 16
        cells.append(((ix + 0.5) / 16 - 0.5, (iy + 0.5) / 16 - 0.5, (iz + 0.5) / 16 - 0.5, h))
    return np.array(cells)


def fit_cells(oracle, cfg, prog, cells):
    out = []
    for c in cells:
        coeffs, err = oracle.oracle_fit(cfg, prog, c[:3] - c[3], c[:3] + c[3], 2, 4)
        out.append(np.concatenate([[err, coeffs[0]], coeffs]))
    return np.array(out).reshape(len(cells), 12)


def worker(rank, world, port, n, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import importlib
    from oracle import hporacle, hpref
    from cases import CASES
    hp = importlib.import_module("hp-adaptive-signed-distance-field-octree_b200")
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    c = CASES["c2_csg"]
    cfg, prog = hpref.make_config(**c["cfg"]), hpref.make_program(c["prog"])
    cells = coarse_cells(n)
    b, e = hp.shard_range(n, rank, world)
    mine = fit_cells(hporacle, cfg, prog, cells[b:e])
    # ragged shards: pad to the largest shard, gather, then cut each rank's piece back to its range
    sizes = [hp.shard_range(n, r, world) for r in range(world)]
    width = max(e2 - b2 for b2, e2 in sizes)
    buf = torch.zeros((width, 12), dtype=torch.float64)
    buf[: e - b] = torch.from_numpy(mine)
    gathered = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(gathered, buf)
    merged = np.concatenate([g.numpy()[: e2 - b2] for g, (b2, e2) in zip(gathered, sizes)])
    ret[rank] = merged
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [37, 64])
def test_sharded_round_equals_single_process(oracle, n):
    from oracle import hpref
    from cases import CASES
    world, port = 2, 29611 + n
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(worker, args=(world, port, n, ret), nprocs=world, join=True)
    c = CASES["c2_csg"]
    single = fit_cells(oracle, hpref.make_config(**c["cfg"]), hpref.make_program(c["prog"]), coarse_cells(n))
    for rank in range(world):
        assert np.array_equal(ret[rank], single), "rank %d holds different records" % rank
