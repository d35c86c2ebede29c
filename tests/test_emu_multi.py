"""The single-call multi-GPU entry points (ps_mctx / ps_mg16_prove / ps_mmsm, csrc/capi_multi.cu) in host
emulation: N "devices" are N contexts with one worker thread each, exchanging through the same code path as
on the GPU box (peer copies degrade to memmove, events to no-ops, the host barriers stay).  Every result must
equal the single-context one bit for bit."""
import os

import pytest

from playsnark_b200 import _lib as L, api, build as B
from tests import parity_cases as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    return L.bind(B.build_host_emulation(os.path.join(ROOT, "tests", "_build")))


@pytest.fixture(scope="module")
def be(lib):
    b = api.Backend(0, lib=lib)
    yield b
    b.close()


@pytest.mark.parametrize("ndev", [1, 2, 3, 8])
def test_mg16_prove_matches_single_device(lib, be, ndev):
    P.multi_groth16_case(lib, be, ndev, log_n=4, seed=ndev)


@pytest.mark.parametrize("ndev,n", [(2, 11), (4, 19)])
def test_mg16_prove_any_gate_count(lib, be, ndev, n):
    P.multi_groth16_case(lib, be, ndev, log_n=0, seed=n, n=n)


def test_mg16_prove_dense_qap(lib, be):
    P.multi_groth16_dense_case(lib, be, 2)


@pytest.mark.parametrize("ndev,group", [(3, L.PS_G1), (4, L.PS_G2)])
def test_mmsm_matches_exponent(lib, ndev, group):
    P.multi_msm_case(lib, ndev, group, 41 if group == L.PS_G1 else 19)


def test_mctx_argument_errors(lib):
    import ctypes as C
    ctx = C.c_void_p()
    assert lib.ps_mctx_create(None, 2, C.byref(ctx)) == L.PS_ERR_ARG
    assert lib.ps_mctx_create((C.c_int * 1)(0), 0, C.byref(ctx)) == L.PS_ERR_ARG
