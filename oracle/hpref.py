"""ctypes bindings of oracle/_ref/libhpref*.so — the reference's own sources compiled here (oracle/Makefile).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs. The product (hp-adaptive-signed-distance-field-octree_b200/) never imports this.
"""
import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class Instr(C.Structure):
    """hpsdf_sdf_instr (include/hpsdf.h)."""
    _fields_ = [("op", C.c_uint32), ("_pad", C.c_uint32), ("handle", C.c_void_p), ("p", C.c_double * 8)]


class Config(C.Structure):
    """hpsdf_config = 80-byte LP64 image of SDF::Config (Include/HP/Config.h:12-43)."""
    _fields_ = [("nearness_type", C.c_uint8), ("_pad0", C.c_uint8 * 7), ("nearness_strength", C.c_double),
                ("continuity_enforce", C.c_uint8), ("_pad1", C.c_uint8 * 7), ("continuity_strength", C.c_double),
                ("enable_logging", C.c_uint8), ("_pad2", C.c_uint8 * 7), ("target_error_threshold", C.c_double),
                ("thread_count", C.c_uint64), ("root_min", C.c_float * 3), ("root_max", C.c_float * 3)]


assert C.sizeof(Config) == 80 and C.sizeof(Instr) == 80

PRIM = dict(sphere=1, box=2, torus=3, capsule=4, plane=5, mesh=16, octree=17)
OP = dict(union=64, intersect=65, subtract=66, negate=67)


def make_config(threshold=1e-10, nearness=0, strength=0.0, continuity=True, cstrength=8.0, threads=1,
                root_min=(-0.5, -0.5, -0.5), root_max=(0.5, 0.5, 0.5), logging=False):
    c = Config()
    C.memset(C.byref(c), 0, 80)
    c.nearness_type = nearness
    c.nearness_strength = strength
    c.continuity_enforce = 1 if continuity else 0
    c.continuity_strength = cstrength
    c.enable_logging = 1 if logging else 0
    c.target_error_threshold = threshold
    c.thread_count = threads
    c.root_min[:] = [np.float32(v) for v in root_min]
    c.root_max[:] = [np.float32(v) for v in root_max]
    return c


def make_program(items):
    """items: list of (opname, params[, handle]) -> ctypes array of Instr."""
    arr = (Instr * len(items))()
    for i, it in enumerate(items):
        name, params = it[0], it[1]
        handle = it[2] if len(it) > 2 else None
        arr[i].op = PRIM.get(name, OP.get(name))
        arr[i].handle = handle
        for k, v in enumerate(params):
            arr[i].p[k] = float(v)
    return arr


def available(fast=False):
    return os.path.exists(os.path.join(HERE, "_ref", "libhpref_fast.so" if fast else "libhpref.so"))


_libs = {}


def lib(fast=False):
    name = "libhpref_fast.so" if fast else "libhpref.so"
    if name in _libs:
        return _libs[name]
    L = C.CDLL(os.path.join(HERE, "_ref", name))
    L.hpref_build.restype = C.c_void_p
    L.hpref_build.argtypes = [C.POINTER(Config), C.POINTER(Instr), C.c_uint32, C.c_int, C.c_uint32, C.c_uint32,
                              C.c_uint32, C.c_double, C.c_int]
    L.hpref_from_block.restype = C.c_void_p
    L.hpref_from_block.argtypes = [C.c_void_p, C.c_size_t]
    L.hpref_destroy.argtypes = [C.c_void_p]
    L.hpref_block_size.restype = C.c_size_t
    L.hpref_block_size.argtypes = [C.c_void_p]
    L.hpref_block_copy.argtypes = [C.c_void_p, C.c_void_p]
    L.hpref_query.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int]
    L.hpref_query_gradient.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int]
    L.hpref_query_ray.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_double, C.c_void_p, C.c_void_p]
    L.hpref_stats.argtypes = [C.c_void_p, C.c_void_p]
    L.hpref_apply_log.argtypes = [C.c_void_p, C.c_void_p]
    L.hpref_fit.restype = C.c_double
    L.hpref_fit.argtypes = [C.POINTER(Config), C.POINTER(Instr), C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32,
                            C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
    L.hpref_fit_chain_batch.argtypes = [C.POINTER(Config), C.POINTER(Instr), C.c_uint32, C.c_size_t, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.hpref_tables.argtypes = [C.c_void_p] * 7
    L.hpref_continuity_triplets.restype = C.c_size_t
    L.hpref_continuity_triplets.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.hpref_mesh_create.restype = C.c_void_p
    L.hpref_mesh_create.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
    L.hpref_mesh_load_obj.restype = C.c_void_p
    L.hpref_mesh_load_obj.argtypes = [C.c_char_p, C.c_int]
    L.hpref_mesh_counts.argtypes = [C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    L.hpref_mesh_arrays.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.hpref_mesh_sdf.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_int]
    L.hpref_mesh_destroy.argtypes = [C.c_void_p]
    L.hpref_hardware_threads.restype = C.c_int
    _libs[name] = L
    return L


STAT_KEYS = ["seconds", "continuity_seconds", "fits", "jobs", "applied_p", "applied_h", "final_total",
             "cg_iterations", "cg_error", "log_size"]


class RefTree:
    """One reference SDF::Octree. mode 0 = literal Octree::Create, mode 1 = deterministic strict-greedy driver."""

    def __init__(self, handle, L, keep=()):
        self.h, self.L, self._keep = handle, L, keep

    @classmethod
    def build(cls, cfg, prog, mode=1, max_degree=11, max_depth=10, total_mode=0, cg_tol=0.0, threads=1, fast=False, mc_seed=None):
        """mc_seed (mode 1): None = exact-mean nearness; an integer = 100 FApprox samples on Philox points (mc_counter)."""
        L = lib(fast)
        L.hpref_set_nearness_mc.argtypes = [C.c_int, C.c_uint64]
        L.hpref_set_nearness_mc(0 if mc_seed is None else 1, 0 if mc_seed is None else mc_seed)
        try:
            h = L.hpref_build(C.byref(cfg), prog, len(prog), mode, max_degree, max_depth, total_mode, cg_tol, threads)
        finally:
            L.hpref_set_nearness_mc(0, 0)
        return cls(h, L, keep=(prog,))

    @classmethod
    def from_block(cls, block, fast=False):
        L = lib(fast)
        buf = np.frombuffer(bytes(block), dtype=np.uint8).copy()
        return cls(L.hpref_from_block(buf.ctypes.data, buf.size), L)

    def block(self):
        n = self.L.hpref_block_size(self.h)
        buf = np.empty(n, dtype=np.uint8)
        self.L.hpref_block_copy(self.h, buf.ctypes.data)
        return buf

    def query(self, pts, threads=1):
        pts = np.ascontiguousarray(pts, dtype=np.float64)
        out = np.empty(len(pts), dtype=np.float64)
        self.L.hpref_query(self.h, pts.ctypes.data, len(pts), out.ctypes.data, threads)
        return out

    def query_gradient(self, pts, threads=1):
        pts = np.ascontiguousarray(pts, dtype=np.float64)
        out = np.empty(len(pts), dtype=np.float64)
        g = np.empty((len(pts), 3), dtype=np.float64)
        self.L.hpref_query_gradient(self.h, pts.ctypes.data, len(pts), out.ctypes.data, g.ctypes.data, threads)
        return out, g

    def query_ray(self, origins, dirs, t_max):
        o = np.ascontiguousarray(origins, np.float64)
        d = np.ascontiguousarray(dirs, np.float64)
        hit = np.zeros(len(o), np.uint8)
        t = np.zeros(len(o), np.float64)
        self.L.hpref_query_ray(self.h, o.ctypes.data, d.ctypes.data, len(o), float(t_max), hit.ctypes.data, t.ctypes.data)
        return hit.astype(bool), t

    def stats(self):
        s = np.zeros(10)
        self.L.hpref_stats(self.h, s.ctypes.data)
        return dict(zip(STAT_KEYS, s.tolist()))

    def apply_log(self):
        n = int(self.stats()["log_size"])
        a = np.zeros((n, 8))
        if n:
            self.L.hpref_apply_log(self.h, a.ctypes.data)
        return a

    def continuity_triplets(self):
        n = self.L.hpref_continuity_triplets(self.h, None, None, None)
        r = np.empty(n, np.int32)
        c = np.empty(n, np.int32)
        v = np.empty(n, np.float64)
        self.L.hpref_continuity_triplets(self.h, r.ctypes.data, c.ctypes.data, v.ctypes.data)
        return r, c, v

    def close(self):
        if self.h:
            self.L.hpref_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()


def ref_fit(cfg, prog, aabb_min, aabb_max, degree, depth, degree_in=0, coeffs_in=None, fast=False):
    """One reference FitPolynomial (Octree.cpp:1007-1093), nearness None. Returns (coeffs[N_degree], raw_err)."""
    L = lib(fast)
    ncoef = NCOEF[degree]
    out = np.zeros(ncoef)
    mn = np.asarray(aabb_min, np.float32)
    mx = np.asarray(aabb_max, np.float32)
    cin = np.ascontiguousarray(coeffs_in, np.float64) if coeffs_in is not None else np.zeros(1)
    err = L.hpref_fit(C.byref(cfg), prog, len(prog), mn.ctypes.data, mx.ctypes.data, degree_in, cin.ctypes.data,
                      degree, depth, out.ctypes.data)
    return out, err


def ref_fit_chain_batch(cfg, prog, aabb_min, aabb_max, d0, d1, depth, threads=1, fast=False, kept_degree=None):
    """Leaf histories replayed with the reference's FitPolynomial (OpenMP over leaves): from scratch at d0[i], then kept-shell
    fits up to d1[i]. Returns (list of coefficient arrays, raw errors of the last fits)."""
    L = lib(fast)
    mn = np.ascontiguousarray(aabb_min, np.float32).reshape(-1, 3)
    mx = np.ascontiguousarray(aabb_max, np.float32).reshape(-1, 3)
    n = len(mn)
    a0, a1, dp = (np.ascontiguousarray(x, np.uint32) for x in (d0, d1, depth))
    out = np.zeros((n, 455))
    err = np.zeros(n)
    kd = np.ascontiguousarray(kept_degree, np.uint32) if kept_degree is not None else None
    L.hpref_fit_chain_batch(C.byref(cfg), prog, len(prog), n, mn.ctypes.data, mx.ctypes.data, a0.ctypes.data, a1.ctypes.data,
                            dp.ctypes.data, kd.ctypes.data if kd is not None else None, out.ctypes.data, err.ctypes.data, threads)
    return [out[i, :NCOEF[int(a1[i])]].copy() for i in range(n)], err


def ref_tables(fast=False):
    L = lib(fast)
    t = dict(nl=np.zeros((13, 11)), counts=np.zeros(13, np.uint32), basis_idx=np.zeros((455, 3), np.uint32),
             recur=np.zeros((13, 2)), roots=np.zeros(2080), weights=np.zeros(2080), face_lookup=np.zeros((3, 4, 2), np.uint32))
    L.hpref_tables(*[t[k].ctypes.data for k in ("nl", "counts", "basis_idx", "recur", "roots", "weights", "face_lookup")])
    return t


class RefMesh:
    def __init__(self, h, L):
        self.h, self.L = h, L

    @classmethod
    def create(cls, verts, tris, bvh=True, fast=False):
        L = lib(fast)
        v = np.ascontiguousarray(verts, np.float32)
        t = np.ascontiguousarray(tris, np.uint32)
        h = L.hpref_mesh_create(v.ctypes.data, len(v), t.ctypes.data, len(t), 1 if bvh else 0)
        return cls(h, L) if h else None

    @classmethod
    def load_obj(cls, path, bvh=True, fast=False):
        L = lib(fast)
        h = L.hpref_mesh_load_obj(path.encode(), 1 if bvh else 0)
        return cls(h, L) if h else None

    def arrays(self):
        nv, nt = C.c_size_t(), C.c_size_t()
        self.L.hpref_mesh_counts(self.h, C.byref(nv), C.byref(nt))
        v = np.empty((nv.value, 3), np.float32)
        t = np.empty((nt.value, 3), np.uint32)
        self.L.hpref_mesh_arrays(self.h, v.ctypes.data, t.ctypes.data)
        return v, t

    def sdf(self, pts, use_bvh=True, threads=1):
        p = np.ascontiguousarray(pts, np.float32)
        out = np.empty(len(p), np.float32)
        self.L.hpref_mesh_sdf(self.h, p.ctypes.data, len(p), out.ctypes.data, 1 if use_bvh else 0, threads)
        return out

    def close(self):
        if self.h:
            self.L.hpref_mesh_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()


def parse_block(block):
    """MemoryBlock (LP64 layout) -> dict of numpy views. Compare field-wise, never memcmp."""
    b = np.frombuffer(bytes(block), dtype=np.uint8)
    ncoef = int(b[:8].view(np.uint64)[0])
    coeffs = b[8:8 + 8 * ncoef].view(np.float64)
    off = 8 + 8 * ncoef
    nnodes = int(b[off:off + 8].view(np.uint64)[0])
    off += 8
    node_dt = np.dtype({"names": ["child", "mn", "mx", "cstart", "deg", "depth"],
                        "formats": ["<u8", ("<f4", 3), ("<f4", 3), "<u8", "u1", "u1"],
                        "offsets": [0, 8, 20, 32, 40, 48], "itemsize": 56})
    nodes = b[off:off + 56 * nnodes].view(node_dt)
    off += 56 * nnodes
    cfg = Config.from_buffer_copy(bytes(b[off:off + 80]))
    assert off + 80 == len(b), "block size mismatch"
    return dict(n_coeffs=ncoef, coeffs=coeffs, n_nodes=nnodes, nodes=nodes, config=cfg)


# LegendreCoeffientCount incl. the reference's truncation quirk at degree 6 (Utility.h:87-106 gives 83, not 84)
NCOEF = [1, 4, 10, 20, 35, 56, 83, 120, 165, 220, 286, 364, 455]


def canonical(block):
    """Canonical structural form: DFS by child slot from the root -> list of (depth, degree, path, coeffs) per leaf,
    independent of node numbering."""
    t = parse_block(block)
    nodes, coeffs = t["nodes"], t["coeffs"]
    leaves = []
    stack = [(0, ())]
    while stack:
        idx, path = stack.pop()
        nd = nodes[idx]
        if nd["child"] == np.uint64(0xFFFFFFFFFFFFFFFF):
            d = int(nd["deg"])
            s = int(nd["cstart"])
            leaves.append((int(nd["depth"]), d, path, coeffs[s:s + NCOEF[d]].copy()))
        else:
            c = int(nd["child"])
            for i in range(7, -1, -1):
                stack.append((c + i, path + (i,)))
    return leaves
