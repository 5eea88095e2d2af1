"""GPU: the reference's own unit tests (Source/Tests/HPUnitTests.cpp:46-316) rewritten against the C++ facade
include/hpsdf.hpp — compiled here with g++ and run against libhpsdf.so."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "hp-adaptive-signed-distance-field-octree_b200")


@pytest.mark.gpu
def test_reference_unit_tests_through_the_cpp_facade(tmp_path):
    exe = str(tmp_path / "facade_test")
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([gxx, "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "facade_test.cpp"),
                           "-L", os.path.join(PKG, "lib"), "-lhpsdf", "-o", exe])
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(PKG, "lib") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    out = subprocess.run([exe], env=env, capture_output=True, text=True, timeout=600)
    print(out.stdout)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "All tests passed!" in out.stdout


def test_facade_header_compiles_without_a_gpu(tmp_path):
    """CPU: the facade is header-only C++17 over the C ABI; it must compile and link against libhpsdf.so here."""
    src = tmp_path / "t.cpp"
    src.write_text('#include "hpsdf.hpp"\nint main() { SDF::Config c; c.IsValid(); SDF::Program p; p.Sphere(0, 0, 0, 0.25); '
                   'static_assert(sizeof(SDF::Config) == 80, "layout"); return hpsdf_device_count() < 0; }\n')
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    exe = str(tmp_path / "t")
    subprocess.check_call([gxx, "-std=c++17", "-I", os.path.join(ROOT, "include"), str(src), "-L", os.path.join(PKG, "lib"),
                           "-lhpsdf", "-o", exe])
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(PKG, "lib"))
    assert subprocess.run([exe], env=env).returncode == 0


@pytest.mark.gpu
def test_function_slice_is_a_batched_query_grid(tmp_path):
    """Octree::OutputFunctionSlice (Octree.cpp:1132-1205) as one batched Query over the slice grid + the reference's colouring."""
    import importlib
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from common import product_cfg
    hp = importlib.import_module("hp-adaptive-signed-distance-field-octree_b200")
    cfg, prog = product_cfg(hp, "sphere_poly_1e8")
    t = hp.Octree()
    t.Create(cfg, prog)
    n = 256
    vals, img = t.OutputFunctionSlice(str(tmp_path / "slice"), 0.1, (-0.5, -0.5, -0.5), (0.5, 0.5, 0.5), n)
    step = np.float32(1.0) / np.float32(n)
    xs = -0.5 + (np.arange(n, dtype=np.float32) * step).astype(np.float64)
    pts = np.stack(list(np.meshgrid(xs, xs, indexing="xy")) + [np.full((n, n), 0.1)], -1).reshape(-1, 3)
    assert np.array_equal(vals.reshape(-1), t.Query(pts))
    truth = np.sqrt((pts[:, 0] - 0.25) ** 2 + pts[:, 1] ** 2 + pts[:, 2] ** 2) - 0.5
    assert np.abs(vals.reshape(-1) - truth).max() < 0.01
    inside = vals <= 1e-6
    assert (img[:, :, 1][inside] == 0).all() and (img[:, :, 2][~inside] == 0).all() and img[:, :, 0].max() == 0
    assert img[:, :, 1][~inside].max() >= 250 and img[:, :, 2][inside].max() >= 250        # each side rescaled to its own range (u8 truncation)
    data = open(str(tmp_path / "slice.bmp"), "rb").read()
    assert data[:2] == b"BM" and len(data) == 54 + 3 * n * n
