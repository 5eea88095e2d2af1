#!/usr/bin/env python
"""bench.py — the reference's headline metric on B200: octree build nodes-fitted/s (+ Query points/s) vs CPU.

    python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (CUDA through the C ABI)
    python bench.py --impl reference [--steps K] [--warmup W]      the reference's own CPU implementation (oracle/_ref)
    torchrun --nproc-per-node N ... bench.py --gpus N ...          N > 1: one rank per GPU (NCCL)

A "step" is one Octree::Create of BASELINE.json configs[1] (C2: union of analytic box + torus + capsule, threshold 1e-8,
continuity off, root [-0.25,0.5]^3). `value` = nodes fitted per second, where "nodes fitted" is the number of
FitPolynomial-equivalents on the strict-greedy path of that build (coarse fits + 9 per applied refinement job; the count
is a constant of the workload, taken from the build's own statistics), so both arms divide the same work by their time.
The timed region of `value` covers the whole Create with the tree left resident in HBM; `e2e` goes through the public
API with host buffers (program/config in, MemoryBlock out, copies inside the timed region).
The same JSON line carries a `query` object: batched Octree::Query points/s on the C2 tree (device-resident points,
inputs larger than L2) with its HBM roofline, and `roofline_frontier`: the fit kernel on the synthetic frontier of
SURVEY.md §8(d) (a real build is too small to fill 148 SMs, so kernel quality is read there).

Prints ONE JSON line on rank 0.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cases import CASES  # noqa: E402

PKG = "hp-adaptive-signed-distance-field-octree_b200"
WORKLOAD = "c2_csg"
METRIC = "octree_build_nodes_fitted_per_s"
QUERY_POINTS = 1 << 24            # 16.7 M points = 512 MB in + 128 MB out: larger than the 126 MB L2
MESH_UV = (1000, 435)             # bumpy torus with 870 000 triangles: the stand-in of configs[2]'s dragon.obj (absent from the reference tree)
MESH_C4_UV = (1000, 800)          # 1.6 M triangles: the stand-in of configs[3]'s Ramesses.obj


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.rows, self.stop, self.index = [], False, index
        self.th = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.th.join(timeout=6)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def product_case(hp, name):
    k = CASES[name]["cfg"]
    cfg = hp.Config(target_error_threshold=k["threshold"], nearness_type=k.get("nearness", 0),
                    nearness_strength=k.get("strength", 0.0), continuity_enforce=1 if k.get("continuity", True) else 0,
                    continuity_strength=k.get("cstrength", 8.0), thread_count=os.cpu_count() or 1,
                    root_min=k.get("root_min", (-0.5,) * 3), root_max=k.get("root_max", (0.5,) * 3))
    return cfg, hp.SdfProgram(CASES[name]["prog"])


def mesh_bench(hp, torch, local, stream, comm, barrier, world, rank, with_cpu):
    """The mesh configs at scale, on procedural stand-ins of the missing dragon.obj / Ramesses.obj:
    C3 = 870 k triangles, threshold 1e-6, continuity strength 8; C4 = 1.6 M triangles, threshold 1e-8, max degree 6, frontier
    sharded over the ranks; C5 = batched Query of 1e9 uniform points on the C3 octree (replicated tree, points sharded).
    Rank 0 at N=1 also times the reference's Mesh::SignedDistanceAtPt on a bounded sample on all host cores."""
    from meshgen import bumpy_torus, mesh_root
    dist = None
    if world > 1:
        import torch.distributed as dist

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.cpu()[0])

    def build(uv, thr, cont, max_degree, reps):
        verts, tris = bumpy_torus(*uv)
        t0 = time.perf_counter()
        mesh = hp.Mesh(verts, tris, device=local)
        create_s = time.perf_counter() - t0
        mn, mx = mesh_root(verts)
        cfg = hp.Config(target_error_threshold=thr, nearness_type=0, nearness_strength=0.0, continuity_enforce=cont,
                        continuity_strength=8.0, thread_count=os.cpu_count() or 1, root_min=mn, root_max=mx)
        prog = hp.SdfProgram([("mesh", [], mesh)])
        opts = hp.BuildOpts(device=local, stream=stream, max_degree=max_degree)
        if comm is not None:
            opts.comm = comm._h
        tree = hp.Octree()
        tree.Create(cfg, prog, opts)                       # warm-up (sample scratch allocation)
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            tree.Create(cfg, prog, opts)
        barrier()
        ms = max_over_ranks(1e3 * (time.perf_counter() - t0) / reps)
        st = tree.stats()
        out = {"triangles": int(len(tris)), "threshold": thr, "continuity": bool(cont), "max_degree": max_degree,
               "create_ms": ms, "fits_per_s": useful_fits(st) / (ms * 1e-3), "fits_evaluated": st["fits_evaluated"],
               "mesh_sdf_evals": st["sdf_evals"], "mesh_sdf_evals_per_s_per_gpu_kernel_time": st["sdf_evals"] / world / (st["fit_kernel_ms"] * 1e-3),
               "mesh_sdf_evals_per_s": st["sdf_evals"] / (ms * 1e-3),
               "fit_and_sample_kernel_ms": st["fit_kernel_ms"], "continuity_ms": st["continuity_ms"], "rounds": st["rounds"],
               "n_nodes": st["n_nodes"], "n_coeffs": st["n_coeffs"], "mesh_upload_and_bvh_s": create_s, "scaling": "strong"}
        return out, tree, mesh, verts, tris, (mn, mx)

    c3, tree3, mesh3, verts3, tris3, (mn, mx) = build(MESH_UV, 1e-6, 1, 11, 3)
    res = {"c3_dragon_standin": c3}

    # ---- C5: 1e9 points on the C3 tree, generated on the device (Philox) in chunks of 2^26, sharded over the ranks ------
    total, chunk = 1_000_000_000, 1 << 26
    mine = total // world + (1 if rank < total % world else 0)
    gen = torch.Generator(device="cuda").manual_seed(0x5DF0C7EE + rank)
    lo = torch.tensor(mn, device="cuda", dtype=torch.float64)
    ext = torch.tensor([mx[i] - mn[i] for i in range(3)], device="cuda", dtype=torch.float64)
    out = torch.empty(chunk, device="cuda", dtype=torch.float64)
    q_ms, done, checksum = 0.0, 0, 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    while done < mine:
        m = min(chunk, mine - done)
        pts = (torch.rand((m, 3), generator=gen, device="cuda", dtype=torch.float64) * ext + lo).contiguous()    # untimed
        e0.record()
        tree3.QueryDevice(pts.data_ptr(), m, out.data_ptr(), stream)
        e1.record()
        torch.cuda.synchronize()
        q_ms += e0.elapsed_time(e1)
        checksum += float(out[:m].sum())
        done += m
        del pts
    q_ms = max_over_ranks(q_ms)
    res["c5_query_1e9"] = {"points": total, "ms": q_ms, "points_per_s": total / (q_ms * 1e-3), "scaling": "strong",
                           "tree": "C3 octree replicated on every rank, points sharded contiguously, generated on the device",
                           "hbm_GBps": total / world * 32 / (q_ms * 1e-3) / 1e9, "checksum_rank0": checksum}
    cpu = None
    if with_cpu and rank == 0:
        try:
            from oracle import hpref, hporacle
            threads = os.cpu_count() or 1
            pts = np.random.default_rng(5).uniform(mn, mx, (20000, 3)).astype(np.float32)
            if hpref.available(fast=True):
                t0 = time.perf_counter()
                rm = hpref.RefMesh.create(verts3, tris3, True, fast=True)
                ref_create = time.perf_counter() - t0
                kind = "reference"
            else:
                t0 = time.perf_counter()
                rm = hporacle.OracleMesh(verts3, tris3)
                ref_create = time.perf_counter() - t0
                kind = "port"
            t0 = time.perf_counter()
            d = rm.sdf(pts, True, threads)
            dt = time.perf_counter() - t0
            # identity is checked against the strict (-ffp-contract=off) restatement, which tests/ pin to the reference built the
            # same way; the -O3 -march=x86-64-v3 reference build timed above contracts FMAs and is 1 ulp off ITSELF in ~9 % of points
            ours = mesh3.SignedDistanceAtPt(pts)
            strict = hporacle.OracleMesh(verts3, tris3).sdf(pts, True, threads)
            cpu = {"mesh_sdf_evals_per_s": len(pts) / dt, "cores": threads, "kind": kind,
                   "sample": "Mesh::SignedDistanceAtPt (BVH) at 20 000 uniform points of the C3 root box",
                   "mesh_setup_s": ref_create, "gpu_bit_identical_to_strict_checker": bool(np.array_equal(ours, strict)),
                   "fma_build_points_differing_from_strict": int((d != strict).sum())}
        except Exception as e:      # the checker is optional here: report, do not fail the bench line
            cpu = {"unavailable": repr(e)}
    if cpu is not None:
        res["cpu_baseline"] = cpu
    del tree3, mesh3
    c4, tree4, mesh4, _, _, _ = build(MESH_C4_UV, 1e-8, 0, 6, 1)
    res["c4_ramesses_standin"] = c4
    return res


def query_flops_per_point(hp, tree):
    """Algorithmic FP64 flops of one Query on `tree`, averaged over uniform points of the root box: per leaf of degree p the
    three Legendre recurrences (about 3 flops per degree and axis: 9p) and the N_p-term sum sum c * Lx * Ly * Lz (4 N_p),
    weighted by the leaf's share of the volume."""
    nodes = hp.parse_block(tree.ToMemoryBlockBytes())["nodes"]
    leaf = nodes["child"] == np.iinfo(np.uint64).max
    if not leaf.any():
        leaf = nodes["deg"] != 13
    ext = (nodes["mx"][leaf] - nodes["mn"][leaf]).astype(np.float64)
    vol = ext.prod(1)
    deg = nodes["deg"][leaf].astype(np.int64)
    ncoef = np.asarray(hp.COEFF_COUNT)[deg]
    return float(((9 * deg + 4 * ncoef) * vol).sum() / vol.sum())


def useful_fits(stats):
    """FitPolynomial-equivalents on the strict-greedy path: 4096 coarse fits + the 9 fits of every applied job."""
    return 4096 + 9 * (stats["jobs_applied_p"] - 4096 + stats["jobs_applied_h"])


def cpu_reference_build(threads, name=WORKLOAD):
    """The reference's own Octree::Create (unmodified sources, oracle/_ref, -O3 + OpenMP build) on the host cores; falls
    back to the C port when oracle/_ref was not built. Returns (seconds, kind, n_nodes)."""
    from oracle import hpref
    c = CASES[name]
    cfg = hpref.make_config(threads=threads, **c["cfg"])
    prog = hpref.make_program(c["prog"])
    if hpref.available(fast=True):
        t0 = time.perf_counter()
        t = hpref.RefTree.build(cfg, prog, mode=0, threads=threads, fast=True)
        dt = time.perf_counter() - t0
        return dt, "reference", t
    from oracle import hporacle
    t0 = time.perf_counter()
    t = hporacle.OracleTree.build(cfg, prog, threads=threads)
    return time.perf_counter() - t0, "port", t


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # the workload's nodes-fitted count is a constant of the deterministic schedule (CPU port, same config)
    from oracle import hporacle, hpref
    c = CASES[WORKLOAD]
    det = hporacle.OracleTree.build(hpref.make_config(threads=threads, **c["cfg"]), hpref.make_program(c["prog"]), threads=threads)
    fits = det.stats()["fits"]
    npts = 4_000_000
    pts = np.random.default_rng(0x5DF0C7EE).uniform(-0.25, 0.5, (npts, 3))
    times, kind, tree = [], "reference", None
    budget_s, t_start = 240.0, time.perf_counter()
    steps_done = 0
    for i in range(args.warmup + args.steps):
        dt, kind, tree = cpu_reference_build(threads)
        if i >= args.warmup:
            times.append(dt)
            steps_done += 1
        elif time.perf_counter() - t_start > 60.0:
            args.warmup = i + 1          # warm-up is bounded too
        if time.perf_counter() - t_start > budget_s and steps_done >= 1:
            break
    ms = 1e3 * float(np.mean(times))
    value = fits / (ms * 1e-3)
    t0 = time.perf_counter()
    tree.query(pts, threads)
    qv = npts / (time.perf_counter() - t0)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "fits/s", "n_gpus": 0, "steps": steps_done,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "nodes_fitted_per_step": fits, "threads": threads},
        "cpu_baseline": {"value": value, "unit": "fits/s", "cores": threads, "kind": kind,
                         "sample": "full Octree::Create of %s per step, %d timed steps" % (WORKLOAD, steps_done)},
        "e2e": {"value": value, "unit": "fits/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "query": {"metric": "query_points_per_s", "value": qv, "unit": "points/s", "points": npts, "threads": threads,
                  "e2e": {"value": qv, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-mesh", action="store_true", help="skip the mesh sub-benchmarks (configs[2..4] stand-ins)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 0)
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    hp = importlib.import_module(PKG)
    dist = None
    comm = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        # NCCL communicator of the library: rank 0's unique id travels over torch.distributed
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(hp.Comm.unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        comm = hp.Comm(bytes(uid.cpu().numpy().tobytes()), rank, world, local)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    stream = torch.cuda.current_stream().cuda_stream
    cfg, prog = product_case(hp, WORKLOAD)
    # closed-form program: fit kernels specialised at run time (NVRTC); the compile happens in the first warm-up step
    opts = hp.BuildOpts(device=local, stream=stream, jit=1)
    if comm is not None:
        opts.comm = comm._h
    fp64_peak = hp.measure_fp64_peak(local, stream)
    hbm_peak, peak_src = peaks()

    # ---- build: value (device-timed, tree left in HBM) -----------------------------------------------------------
    tree = hp.Octree()
    jit_note = "fit kernels specialised to the SDF program at run time (NVRTC, compiled during warm-up)"
    try:
        tree.Create(cfg, prog, opts)
    except hp.HpsdfError as e:
        # no NVRTC on this box: the interpreted kernels are the same arithmetic, ~1.8x slower in the fit launches
        jit_note = "unavailable (%s): interpreted fit kernels" % e
        opts.jit = 2
    for _ in range(args.warmup):
        tree.Create(cfg, prog, opts)
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    agg = dict(fit_ms=0.0, flops=0.0, launches=0, fits_eval=0, rounds=0, replay_ms=0.0)
    with ClockSampler(local) as clk:
        barrier()
        t_wall0 = time.perf_counter()
        for i in range(args.steps):
            ev[i][0].record()
            tree.Create(cfg, prog, opts)
            ev[i][1].record()
            s = tree.stats()
            agg["fit_ms"] += s["fit_kernel_ms"]; agg["flops"] += s["algorithmic_flops"]; agg["launches"] += s["kernel_launches"]
            agg["fits_eval"] += s["fits_evaluated"]; agg["rounds"] += s["rounds"]; agg["replay_ms"] += s["host_replay_ms"]
        barrier()
        wall_ms = 1e3 * (time.perf_counter() - t_wall0)
        dev_ms = sum(a.elapsed_time(b) for a, b in ev)
        # ---- query: device-resident points, inputs larger than L2 ------------------------------------------------
        n_q = QUERY_POINTS
        gen = torch.Generator(device="cuda").manual_seed(0x5DF0C7EE + rank)
        pts = (torch.rand((n_q, 3), generator=gen, device="cuda", dtype=torch.float64) * 0.75 - 0.25).contiguous()
        out = torch.empty(n_q, device="cuda", dtype=torch.float64)
        for _ in range(3):
            tree.QueryDevice(pts.data_ptr(), n_q, out.data_ptr(), stream)
        barrier()
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        q_reps = 10
        q0.record()
        for _ in range(q_reps):
            tree.QueryDevice(pts.data_ptr(), n_q, out.data_ptr(), stream)
        q1.record()
        barrier()
        q_ms = q0.elapsed_time(q1) / q_reps
    stats = tree.stats()
    fits = useful_fits(stats)
    t = torch.tensor([dev_ms, wall_ms, q_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms, q_ms = [float(x) for x in t.cpu()]
    ms_per_step = max(dev_ms, wall_ms) / args.steps if world > 1 else dev_ms / args.steps
    value = fits / (ms_per_step * 1e-3)                    # one tree built cooperatively by all ranks: strong scaling
    q_value = world * n_q / (q_ms * 1e-3)                  # every rank queries its own batch on the replicated tree: weak
    q_flops = query_flops_per_point(hp, tree)

    # ---- e2e through the public API with host buffers -------------------------------------------------------------
    barrier()
    e2e_times, blk_bytes = [], 0
    for i in range(max(3, min(args.steps, 10))):
        t0 = time.perf_counter()
        t2 = hp.Octree()
        t2.Create(cfg, prog, opts)
        blk = t2.ToMemoryBlock()                           # device -> host copy of the whole tree (malloc-owned MemoryBlock)
        torch.cuda.synchronize()
        e2e_times.append(time.perf_counter() - t0)
        blk_bytes = blk.size
        blk.free()
        t2.Clear()
    e2e_ms = 1e3 * float(np.mean(e2e_times[1:]))
    te = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms = float(te.cpu()[0])
    h2d = 80 + 80 * len(CASES[WORKLOAD]["prog"]) + 32 * stats["jobs_evaluated"]      # config + program + one 32-byte record per refinement job
    d2h = blk_bytes + 16 * stats["fits_evaluated"]                                    # MemoryBlock + {error, c0} records
    # query e2e: pinned host points in, host values out
    n_qe = 1 << 22
    hpts = torch.empty((n_qe, 3), dtype=torch.float64).pin_memory()
    hpts.copy_(pts[:n_qe].cpu())
    hout = torch.empty(n_qe, dtype=torch.float64).pin_memory()
    import ctypes
    for i in range(4):
        if i == 1:
            barrier(); tq0 = time.perf_counter()
        hp._check(hp.lib().hpsdf_query(tree._h, ctypes.c_void_p(hpts.data_ptr()), n_qe, ctypes.c_void_p(hout.data_ptr())))
    qe_s = (time.perf_counter() - tq0) / 3
    qe_value = world * n_qe / qe_s

    # ---- synthetic frontier: all 32768 cells of a depth-5 grid as jobs at p = 2..4 -------------------
    frontier = {}
    if rank == 0:
        hp.set_jit(opts.jit == 1)
        for p in (2, 3, 4):
            fb = hp.bench_frontier(cfg, prog, 5, p, repeats=3, device=local, stream=stream)
            frontier["p%d" % p] = {"ms": fb["ms_per_launch"], "jobs": fb["jobs"], "fits_per_s": fb["fits"] / (fb["ms_per_launch"] * 1e-3),
                                   "sdf_evals_per_s": fb["sdf_evals"] / (fb["ms_per_launch"] * 1e-3),
                                   "algorithmic_tflops": fb["algorithmic_flops"] / (fb["ms_per_launch"] * 1e-3) / 1e12,
                                   "frac_of_fp64_peak": fb["algorithmic_flops"] / (fb["ms_per_launch"] * 1e-3) / 1e12 / fp64_peak}
        hp.set_jit(False)
    mesh_line = None
    if not args.no_mesh:
        mesh_line = mesh_bench(hp, torch, local, stream, comm, barrier, world, rank, world == 1 and not args.no_cpu_baseline)

    if rank != 0:
        if comm is not None:
            comm.close()
        if dist is not None:
            dist.destroy_process_group()
        return

    fit_tflops = agg["flops"] / max(agg["fit_ms"] * 1e-3, 1e-12) / 1e12
    line = {
        "metric": METRIC, "value": value, "unit": "fits/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "nodes_fitted_per_step": fits, "fits_evaluated_per_step": stats["fits_evaluated"],
                   "rounds_per_step": stats["rounds"], "n_nodes": stats["n_nodes"], "n_coeffs": stats["n_coeffs"],
                   "multi_gpu": "one tree built by all ranks: mesh / octree programs and rounds of >= 2^18 fits are sharded (grouped NCCL broadcasts "
                                "per round); small closed-form rounds are evaluated redundantly on every rank (an exchange costs more than the "
                                "kernels), so this 4 ms host-bound build does not speed up with N — mesh_build and query are the paths that scale",
                   "jit": jit_note,
                   "l2": "build inputs are a 480-byte program; query inputs (%d MB) are larger than L2" % (n_q * 32 >> 20),
                   "timing": "CUDA events on the build stream around each Create (max over ranks)"},
        "clocks": clk.summary(),
        "e2e": {"value": fits / (e2e_ms * 1e-3), "unit": "fits/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": int(agg["launches"]),
        "roofline": {"kernel": "fitKernel<D> specialised to the program (all fit launches of the timed builds)", "bound": "fp64",
                     "achieved": fit_tflops, "peak": fp64_peak, "unit": "TFLOP/s", "frac": fit_tflops / fp64_peak,
                     "traffic": None,
                     "traffic_note": "the fit kernels read a 32-byte task and write N_d coefficients per fit; ncu on the frontier launch "
                                     "(262 144 fits @2): 8.4 MB DRAM read, writes stay in L2 (profiles/r1_fit_kernel_jit.md)",
                     "note": "achieved = SURVEY.md 8d algorithmic FLOPs (sum-factorised contraction + c_F = %.0f per SDF sample, sqrt/div "
                             "counted as 1) / device time of the fit launches; peak = DFMA rate measured in this run "
                             "(hpsdf_measure_fp64_peak; MEASURED_PEAKS.json holds no FP64 figure); a C2 build has only ~%d fits per "
                             "round, see roofline_frontier for full-GPU launches" % (stats["sdf_flops_per_eval"], stats["fits_evaluated"] // max(stats["rounds"], 1)),
                     "fit_kernel_ms_per_step": agg["fit_ms"] / args.steps, "host_replay_ms_per_step": agg["replay_ms"] / args.steps},
        "roofline_frontier": frontier,
        "query": {"metric": "query_points_per_s", "value": q_value, "unit": "points/s", "points_per_gpu": n_q, "ms": q_ms,
                  "scaling": "weak",
                  "e2e": {"value": qe_value, "unit": "points/s", "h2d_bytes_per_step": n_qe * 24, "d2h_bytes_per_step": n_qe * 8},
                  "roofline": {"kernel": "queryKernel", "bound": "hbm", "achieved": n_q * 32 / (q_ms * 1e-3) / 1e9,
                               "peak": hbm_peak, "unit": "GB/s", "frac": n_q * 32 / (q_ms * 1e-3) / 1e9 / hbm_peak,
                               "traffic": 525.06e6 if n_q == (1 << 24) else None,
                               "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one launch on 2^24 points, profiles/r1_query_kernel.md "
                                                 "(algorithmic bytes of the same launch: 536.9e6)",
                               "peak_source": peak_src,
                               "fp64": {"algorithmic_flops_per_point": q_flops, "achieved": q_flops * n_q / (q_ms * 1e-3) / 1e12,
                                        "peak": fp64_peak, "unit": "TFLOP/s", "frac": q_flops * n_q / (q_ms * 1e-3) / 1e12 / fp64_peak,
                                        "note": "the kernel is neither HBM- nor FP64-bound: ncu shows 19 cycles of long-scoreboard stall per "
                                                "issued instruction (dependent node and coefficient gathers from L2), issue slots 33 % busy, "
                                                "FP64 pipe 22 % (profiles/r1_query_kernel.md)"}}},
    }
    if mesh_line is not None:
        line["mesh_build"] = mesh_line
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        dt, kind, ctree = cpu_reference_build(threads)
        cq = np.random.default_rng(1).uniform(-0.25, 0.5, (2_000_000, 3))
        t0 = time.perf_counter()
        ctree.query(cq, threads)
        cqv = len(cq) / (time.perf_counter() - t0)
        line["cpu_baseline"] = {"value": fits / dt, "unit": "fits/s", "cores": threads, "kind": kind,
                                "sample": "one full Octree::Create of %s (%.1f s); query: 2e6 points, %d OpenMP threads" % (WORKLOAD, dt, threads),
                                "query_points_per_s": cqv}
    print(json.dumps(line))
    if comm is not None:
        comm.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
