"""The complete C ABI, compiled with -DPS_HOST_EMU (kernel bodies driven by serial loops on the CPU),
against the oracle.  Covers the host orchestration, index arithmetic and data formats that the GPU
suite (test_gpu_parity.py) re-checks on the device through the product library."""
import os

import pytest

from playsnark_b200 import _lib as L, api, build as B
from tests import parity_cases as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def be():
    so = B.build_host_emulation(os.path.join(ROOT, "tests", "_build"))
    lib = L.bind(so)
    assert b"HOST EMULATION" in lib.ps_version()
    b = api.Backend(0, lib=lib)
    yield b
    b.close()


def test_codec(be): P.codec_roundtrip(be, 6)
def test_msm_golden(be): P.msm_golden(be)
def test_msm_errors(be): P.msm_errors(be)
def test_msm_linearity(be): P.msm_linearity(be, 40)


@pytest.mark.parametrize("kind", ["rand", "ones", "neg", "small", "zero", "edge"])
def test_msm_g1_kinds(be, kind): P.msm_exponent_check(be, L.PS_G1, 41, kind)


@pytest.mark.parametrize("n", [1, 2, 3, 100, 257])
def test_msm_g1_sizes(be, n): P.msm_exponent_check(be, L.PS_G1, n)


@pytest.mark.parametrize("wb,tables", [(4, 1), (5, 3), (7, 100), (9, 2), (13, 1)])
def test_msm_g1_windows(be, wb, tables): P.msm_exponent_check(be, L.PS_G1, 50, "rand", wb, tables)


def test_msm_g1_skewed_large(be): P.msm_exponent_check(be, L.PS_G1, 600, "ones")
def test_msm_g1_vs_naive(be): P.msm_vs_naive(be, L.PS_G1, 9)
def test_msm_g2_vs_naive(be): P.msm_vs_naive(be, L.PS_G2, 5)


@pytest.mark.parametrize("kind", ["rand", "small", "edge"])
def test_msm_g2(be, kind): P.msm_exponent_check(be, L.PS_G2, 19, kind)


def test_msm_g2_tables(be): P.msm_exponent_check(be, L.PS_G2, 19, "rand", 6, 4)
def test_ntt(be): P.ntt_cases(be, 7)
def test_ntt_properties(be): P.ntt_properties(be, 9)
def test_readme_quotient(be): P.readme_quotient(be)


@pytest.mark.parametrize("n", [2, 3, 5, 8, 13, 16, 33])
def test_quotient_chain(be, n): P.quotient_vs_div2(be, n, seed=n)


def test_quotient_mixed(be): P.quotient_vs_div2(be, 12, seed=5, circuit="mixed")
def test_readme_groth16(be): P.readme_groth16(be)
def test_readme_phgr13(be): P.readme_phgr13(be)
def test_groth16_mixed(be): P.groth16_circuit(be, 10, seed=3)
def test_groth16_chain_negative_witness(be): P.groth16_circuit(be, 8, seed=4, circuit="chain", verify=False)
def test_phgr13_mixed(be): P.phgr13_circuit(be, 9, seed=6)


@pytest.mark.parametrize("n", [4, 16, 2, 3, 7, 13])
def test_sparse_quotient(be, n):
    # the reference accepts any gate count; 3, 7, 13: interpolation tree over the next power of two with dummy leaves
    P.sparse_quotient_vs_dense(be, n, seed=n)


def test_sparse_groth16_exponent_check(be): P.groth16_sparse_exponent_check(be, 5, seed=8)


def test_config_c2_shape_small(be): P.config_c2(be, n=16)


def test_readme_flow_through_api(be): P.readme_flow_through_api(be)


def test_sparse_phgr13_exponent_check(be): P.phgr13_sparse_exponent_check(be, 4, seed=11)


@pytest.mark.parametrize("n", [8, 5])
def test_device_setups_vs_oracle(be, n): P.device_setups_vs_oracle(be, n, seed=n)


def test_device_setups_long_rows(be, monkeypatch):
    """variables that occur in many gates: their transposed-SpMV sums are cut into segments (here of 2 entries)"""
    monkeypatch.setenv("PLAYSNARK_B200_SPMVT_SEG", "2")
    P.device_setups_vs_oracle(be, 8, seed=31)


@pytest.mark.parametrize("parts,world", [(1, 2), (2, 3), (4, 5)])
def test_sharded_steps_recombine(be, parts, world): P.sharded_steps_recombine(be, 4, parts, world, seed=21 + parts)


@pytest.mark.parametrize("n,parts,world", [(11, 2, 3), (21, 4, 4)])
def test_sharded_steps_recombine_any_n(be, n, parts, world): P.sharded_steps_recombine(be, 0, parts, world, seed=n, n=n)


@pytest.mark.parametrize("wb,tables,kind", [(16, 16, "rand"), (16, 1, "edge"), (20, -1, "rand"), (17, 3, "ones")])
def test_msm_two_pass_scatter(be, wb, tables, kind):
    """the partitioned two-pass scatter of the counting sort (forced on: option msm_scatter = 2)"""
    be.set_option("msm_scatter", 2)
    try:
        P.msm_exponent_check(be, L.PS_G1, 300, kind, wb, tables)
        P.msm_exponent_check(be, L.PS_G2, 40, kind, wb, tables)
    finally:
        be.set_option("msm_scatter", 1)


def test_pairing_checks(be): P.pairing_checks(be)
def test_groth16_verify(be): P.groth16_verify_cases(be)
def test_phgr13_verify(be): P.phgr13_verify_cases(be)
def test_verify_device_setup_flow(be): P.verify_device_setup_flow(be, n=8)

