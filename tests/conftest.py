import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def oracle_lib():
    """The plain-C restatement (built on demand; gcc is in the image)."""
    from oracle import bindings
    bindings.build("port")
    return bindings.OracleLib()


@pytest.fixture(scope="session")
def ref_lib():
    """The reference's own code (oracle/_ref); only where it was built / travelled."""
    from oracle import bindings
    if os.path.isdir("/root/reference/src"):
        bindings.build("ref")
    if not bindings.have_ref():
        pytest.skip("oracle/_ref/libgaselect_ref.so not available")
    return bindings.RefLib()


@pytest.fixture(scope="session")
def engine():
    """The product library (gaselect_b200/libgaselect_b200.so), built on demand with nvcc."""
    import subprocess
    from gaselect_b200 import _capi
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "gaselect_b200", "csrc")], check=True)
    return _capi.lib()


def cuda_device_count():
    import ctypes
    try:
        rt = ctypes.CDLL("libcuda.so.1")
    except OSError:
        return 0
    if rt.cuInit(0) != 0:
        return 0
    n = ctypes.c_int(0)
    rt.cuDeviceGetCount(ctypes.byref(n))
    return n.value
