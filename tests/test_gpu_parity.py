"""Parity tests proper: the CUDA path, called through the product C ABI (libplaysnark_b200.so),
against the oracle / golden fixtures, plus size-independent properties at BASELINE.json's sizes."""

import pytest

from playsnark_b200 import _lib as L, api
from tests import parity_cases as P

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def be():
    lib = L.load()
    assert b"sm_100a" in lib.ps_version()      # the product build, not the host emulation
    b = api.Backend(0)
    before = b.launch_count()
    yield b
    assert b.launch_count() > before           # kernels really ran
    b.close()


def test_codec(be): P.codec_roundtrip(be, 40)
def test_msm_golden(be): P.msm_golden(be)
def test_msm_errors(be): P.msm_errors(be)


@pytest.mark.parametrize("kind", ["rand", "ones", "neg", "small", "zero", "edge"])
def test_msm_g1_kinds(be, kind): P.msm_exponent_check(be, L.PS_G1, 1000, kind)


@pytest.mark.parametrize("n", [1, 2, 3, 31, 257, 4096, 1 << 14])
def test_msm_g1_sizes(be, n): P.msm_exponent_check(be, L.PS_G1, n)


@pytest.mark.parametrize("wb,tables", [(4, 1), (5, 3), (7, 100), (9, 2), (13, 1), (16, 1), (16, 16)])
def test_msm_g1_windows(be, wb, tables): P.msm_exponent_check(be, L.PS_G1, 3000, "rand", wb, tables)


def test_msm_g1_skewed_large(be): P.msm_exponent_check(be, L.PS_G1, 1 << 16, "ones")
def test_msm_g1_small_scalars_large(be): P.msm_exponent_check(be, L.PS_G1, 1 << 16, "small")
def test_msm_g1_vs_naive(be): P.msm_vs_naive(be, L.PS_G1, 48)
def test_msm_g2_vs_naive(be): P.msm_vs_naive(be, L.PS_G2, 24)


@pytest.mark.parametrize("kind", ["rand", "small", "edge", "ones"])
def test_msm_g2(be, kind): P.msm_exponent_check(be, L.PS_G2, 700, kind)


@pytest.mark.parametrize("n", [1, 5, 4096])
def test_msm_g2_sizes(be, n): P.msm_exponent_check(be, L.PS_G2, n)


def test_msm_g2_tables(be): P.msm_exponent_check(be, L.PS_G2, 500, "rand", 8, 4)


def test_msm_large_g1(be):
    # config C4 scale: 2^20 points, checked in the exponent
    P.msm_exponent_check(be, L.PS_G1, 1 << 20, "rand")


def test_msm_large_g2(be):
    P.msm_exponent_check(be, L.PS_G2, 1 << 17, "rand")


def test_msm_g1_2p24_bench_configuration(be):
    # BASELINE configs[3] at the headline size, exactly as bench.py runs it: all window tables, automatic
    # window (c = 22), automatic two-pass scatter (entry array > L2)
    c, W, T = P.msm_exponent_check_big(be, L.PS_G1, 24)
    assert W == T and c >= 20


def test_msm_g2_2p20_bench_configuration(be):
    c, W, T = P.msm_exponent_check_big(be, L.PS_G2, 20)
    assert W == T


def test_msm_g2_2p22(be):
    P.msm_exponent_check_big(be, L.PS_G2, 22, resident=False)


def test_msm_linearity(be): P.msm_linearity(be, 5000)


def test_ntt(be): P.ntt_cases(be, 10)


@pytest.mark.parametrize("log_n", [11, 16, 20])
def test_ntt_properties(be, log_n): P.ntt_properties(be, log_n)


def test_readme_quotient(be): P.readme_quotient(be)


@pytest.mark.parametrize("n", [2, 3, 5, 8, 13, 16, 33, 64, 100])
def test_quotient_chain(be, n): P.quotient_vs_div2(be, n, seed=n)


def test_quotient_mixed(be): P.quotient_vs_div2(be, 48, seed=5, circuit="mixed")
def test_readme_groth16(be): P.readme_groth16(be)
def test_readme_phgr13(be): P.readme_phgr13(be)
def test_groth16_mixed(be): P.groth16_circuit(be, 24, seed=3)
def test_groth16_chain_negative_witness(be): P.groth16_circuit(be, 32, seed=4, circuit="chain", verify=False)
def test_phgr13_mixed(be): P.phgr13_circuit(be, 20, seed=6)


@pytest.mark.parametrize("n", [4, 16, 64, 2, 3, 13, 100])
def test_sparse_quotient(be, n): P.sparse_quotient_vs_dense(be, n, seed=n)


@pytest.mark.parametrize("log_n", [8, 12, 16, 20])
def test_sparse_groth16_exponent_check(be, log_n):
    # config C3: 2^16 constraints, full prove (interpolation + NTT quotient + 3 MSMs), exponent-level parity
    P.groth16_sparse_exponent_check(be, log_n, seed=log_n)


@pytest.mark.parametrize("log_n", [6, 12, 16])
def test_sparse_phgr13_exponent_check(be, log_n): P.phgr13_sparse_exponent_check(be, log_n, seed=log_n)


@pytest.mark.parametrize("n", [1000, 100000])
def test_sparse_any_gate_count(be, n):
    # the reference takes any number of gates: tree over the next power of two with dummy leaves (fused low levels included)
    P.groth16_sparse_exponent_check(be, 0, seed=n % 97, n=n)
    if n <= 1000:
        P.phgr13_sparse_exponent_check(be, 0, seed=n % 89, n=n)


def test_config_c2_dense_2p10(be):
    # BASELINE configs[1]: 2^10 multiplication gates, dense QAP (101 MB), Groth16 + PHGR13
    P.config_c2(be)


def test_readme_flow_through_api(be): P.readme_flow_through_api(be)


@pytest.mark.parametrize("n", [16, 13, 64])
def test_device_setups_vs_oracle(be, n): P.device_setups_vs_oracle(be, n, seed=n)


@pytest.mark.parametrize("log_n,parts,world", [(6, 1, 2), (10, 4, 8), (14, 2, 3)])
def test_sharded_steps_recombine(be, log_n, parts, world):
    # every entry point of the multi-GPU Groth16 flow, recombined on one GPU against ps_g16_prove
    P.sharded_steps_recombine(be, log_n, parts, world, seed=log_n, device="cuda")


@pytest.mark.parametrize("n,wb,tables,kind", [(5000, 16, 16, "rand"), (1 << 16, 0, -1, "rand"), (1 << 18, 20, -1, "small"),
                                              (3000, 17, 3, "ones"), (1 << 21, 0, -1, "rand")])
def test_msm_two_pass_scatter(be, n, wb, tables, kind):
    # block-cooperative staging pass + L2-resident fine scatter, forced on; same bytes as the one-pass sort
    be.set_option("msm_scatter", 2)
    try:
        P.msm_exponent_check(be, L.PS_G1, n, kind, wb, tables)
    finally:
        be.set_option("msm_scatter", 1)


def test_pairing_checks(be): P.pairing_checks(be)
def test_groth16_verify(be): P.groth16_verify_cases(be, n=16)
def test_phgr13_verify(be): P.phgr13_verify_cases(be, n=12)
def test_verify_device_setup_flow(be): P.verify_device_setup_flow(be, n=1 << 10)


def test_no_device_is_loud():
    lib = L.load()
    import ctypes as C
    ctx = C.c_void_p()
    assert lib.ps_ctx_create(9999, C.byref(ctx)) == L.PS_ERR_CUDA
