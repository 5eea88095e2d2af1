import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def hp():
    """The product package (ctypes over lib/libhpsdf.so); built on demand (nvcc cross-compiles without a GPU)."""
    import importlib
    import subprocess
    mod = importlib.import_module("hp-adaptive-signed-distance-field-octree_b200")
    if not os.path.exists(mod.LIB_PATH) and "HPSDF_LIB" not in os.environ:
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "hp-adaptive-signed-distance-field-octree_b200"), "-j8"],
                              stdout=subprocess.DEVNULL)
    return mod


@pytest.fixture(scope="session")
def oracle():
    """The plain-C restatement (oracle/hp_oracle.c), compiled on demand with gcc."""
    from oracle import hporacle
    hporacle.lib()
    return hporacle


@pytest.fixture(scope="session")
def ref():
    """The reference's own sources compiled here (oracle/_ref/libhpref.so); skipped where it was not built."""
    from oracle import hpref
    if not hpref.available():
        pytest.skip("oracle/_ref/libhpref.so not built (needs /root/reference at build time)")
    return hpref
