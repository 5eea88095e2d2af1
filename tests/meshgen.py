"""Procedural closed meshes standing in for the reference's missing dragon.obj / Ramesses.obj."""
import numpy as np


def bumpy_torus(U=60, V=40):
    """Watertight 'bumpy torus' grid: R = 0.30 + 0.03 sin(5u) cos(3v), r = 0.10 + 0.02 sin(7v + 2u).
    2 U V triangles, counter-clockwise seen from outside. U=1000, V=435 is the 870 k-triangle dragon stand-in,
    U=1000, V=800 the 1.6 M-triangle Ramesses stand-in."""
    u = (np.arange(U) * (2 * np.pi / U))[:, None]
    v = (np.arange(V) * (2 * np.pi / V))[None, :]
    R = 0.30 + 0.03 * np.sin(5 * u) * np.cos(3 * v)
    r = 0.10 + 0.02 * np.sin(7 * v + 2 * u)
    x = (R + r * np.cos(v)) * np.cos(u)
    y = (R + r * np.cos(v)) * np.sin(u)
    z = r * np.sin(v) + 0 * u
    verts = np.stack([x, y, z], -1).reshape(-1, 3).astype(np.float32)
    i, j = np.meshgrid(np.arange(U), np.arange(V), indexing="ij")
    a = (i * V + j).ravel()
    b = (((i + 1) % U) * V + j).ravel()
    c = (((i + 1) % U) * V + (j + 1) % V).ravel()
    d = (i * V + (j + 1) % V).ravel()
    tris = np.concatenate([np.stack([a, b, c], 1), np.stack([a, c, d], 1)]).astype(np.uint32)
    return verts, tris


def mesh_root(verts):
    """Root box of the mesh configs: AABB centre +- 0.6 * max extent."""
    mn, mx = verts.min(0).astype(np.float64), verts.max(0).astype(np.float64)
    c, e = (mn + mx) / 2, (mx - mn).max()
    return tuple((c - 0.6 * e).tolist()), tuple((c + 0.6 * e).tolist())
