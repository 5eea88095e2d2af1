"""CPU: the C-ABI library loads and exports every symbol include/hpsdf.h declares; host-only entry points work; compute
entry points fail loudly (no CPU fallback) when there is no GPU. No compute calls are made here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "hpsdf.h")).read()
    return sorted(set(re.findall(r"HPSDF_API[^;(]*?\b(hpsdf_\w+)\s*\(", text)))


def test_header_declares_what_the_binding_lists(hp):
    assert declared_symbols() == sorted(hp.EXPORTS)


def test_library_exports_every_declared_symbol(hp):
    L = C.CDLL(hp.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(L, s), "libhpsdf.so does not export %s" % s
    assert b"sm_100a" in hp.lib().hpsdf_version()


def test_struct_sizes_match_the_reference_layout(hp):
    assert C.sizeof(hp.Config) == 80                      # sizeof(SDF::Config) on LP64
    assert hp.Config.nearness_strength.offset == 8 and hp.Config.continuity_enforce.offset == 16
    assert hp.Config.continuity_strength.offset == 24 and hp.Config.enable_logging.offset == 32
    assert hp.Config.target_error_threshold.offset == 40 and hp.Config.thread_count.offset == 48
    assert hp.Config.root_min.offset == 56 and hp.Config.root_max.offset == 68
    assert C.sizeof(hp.Instr) == 80


def test_config_defaults_and_validation(hp):
    c = hp.Config()                                       # Config::Config (Config.cpp:5-14)
    assert c.target_error_threshold == 1e-10 and c.nearness_type == hp.NEARNESS_NONE
    assert c.continuity_enforce == 1 and c.continuity_strength == 8.0 and c.thread_count >= 1
    assert list(c.root_min) == [-0.5] * 3 and list(c.root_max) == [0.5] * 3 and c.enable_logging == 0
    c.IsValid()
    for bad in (dict(target_error_threshold=0.0), dict(thread_count=0), dict(root_max=(-0.5, 0.5, 0.5)),
                dict(nearness_type=hp.NEARNESS_EXPONENTIAL, nearness_strength=0.0), dict(continuity_strength=-1.0)):
        with pytest.raises(hp.HpsdfError) as e:
            hp.Config(**bad).IsValid()
        assert e.value.status == hp.ERR_INVALID_ARG
    o = hp.BuildOpts()
    assert o.max_degree == 11 and o.max_depth == 10 and o.total_mode == hp.TOTAL_REFERENCE


def test_shard_range_partitions_exactly(hp):
    for n in (0, 1, 7, 4096, 36865):
        for world in (1, 2, 3, 8):
            parts = [hp.shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in parts]
            assert max(sizes) - min(sizes) <= 1


def test_status_strings(hp):
    L = hp.lib()
    assert L.hpsdf_status_string(0) == b"ok"
    assert b"no CPU path" in L.hpsdf_status_string(hp.ERR_NO_DEVICE)


def test_compute_fails_loudly_without_a_gpu(hp):
    if hp.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(hp.HpsdfError) as e:
        hp.Octree().Create(hp.Config(), hp.SdfProgram([("sphere", [0, 0, 0, 0.25])]))
    assert e.value.status == hp.ERR_NO_DEVICE
    with pytest.raises(hp.HpsdfError):
        hp.SdfProgram([("sphere", [0, 0, 0, 0.25])]).eval(np.zeros((4, 3)))
    t = hp.Octree()
    with pytest.raises(hp.HpsdfError):
        t.FromMemoryBlock(hp.MemoryBlock.frombytes(b"\0" * 200))
    with pytest.raises(hp.HpsdfError):
        t.Query(np.zeros((1, 3)))
    with pytest.raises(TypeError):
        t.Create(hp.Config(), lambda p: 0.0)              # host lambdas stay on the reference's CPU path


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the product package may import, link or call it."""
    pkg = os.path.join(ROOT, "hp-adaptive-signed-distance-field-octree_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or f == "Makefile":
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+\.*oracle", text, re.M), f
                assert "hp_oracle" not in text and "hpref" not in text and "libhporacle" not in text, f
                assert not re.search(r'#include\s+"[^"]*oracle', text), f


@pytest.mark.parametrize("degree", [1, 2, 7, 11])
def test_jit_specialisation_builds_for_sm100a_without_a_gpu(hp, degree):
    """csrc/jit.cpp: generated straight-line SDF + fit_kernel_body.cuh compile with NVRTC (headers embedded in the library)."""
    from cases import CSG_C2
    src, nbytes = hp.jit_compile_check(hp.SdfProgram(CSG_C2), degree)
    assert "sdfEval" in src and "0x1." in src          # parameters as exact hex-float literals
    assert nbytes > 10000


def test_jit_rejects_mesh_and_octree_programs(hp):
    prog = hp.SdfProgram([("sphere", [0, 0, 0, 0.3])])
    prog._instr[0].op = hp.PRIM["octree"]             # no handle: resolveProgram refuses before the generator is reached
    with pytest.raises(Exception):
        hp.jit_compile_check(prog, 2)
