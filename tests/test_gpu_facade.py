"""GPU: the reference's own unit tests (Source/Tests/HPUnitTests.cpp:46-316) rewritten against the C++ facade
include/hpsdf.hpp — compiled here with g++ and run against libhpsdf.so."""
import os
import subprocess

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
