import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def host_check():
    """Host build of the device field/curve templates (tests/host_check.cpp)."""
    import ctypes
    bdir = os.path.join(ROOT, "tests", "_build")
    os.makedirs(bdir, exist_ok=True)
    so = os.path.join(bdir, "libhost_check.so")
    srcs = [os.path.join(ROOT, "tests", "host_check.cpp"),
            os.path.join(ROOT, "playsnark_b200", "csrc", "field.cuh"),
            os.path.join(ROOT, "playsnark_b200", "csrc", "curve.cuh"),
            os.path.join(ROOT, "playsnark_b200", "csrc", "constants.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, srcs[0]])
    return ctypes.CDLL(so)
