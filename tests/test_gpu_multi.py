"""Multi-GPU paths on real devices (skipped where the box has fewer GPUs than the case needs):
  - ps_mctx / ps_mg16_prove / ps_mmsm: one host call, peer copies over NVLink, vs the single-device result;
  - playsnark_b200.dist over torch.distributed + NCCL with one process per GPU (torchrun), vs the same."""
import os
import subprocess
import sys

import pytest

from playsnark_b200 import _lib as L, api
from tests import parity_cases as P

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def gpu_count():
    import torch
    return torch.cuda.device_count()


@pytest.fixture(scope="module")
def be():
    b = api.Backend(0)
    yield b
    b.close()


@pytest.mark.parametrize("ndev,log_n", [(1, 10), (2, 12), (2, 16), (4, 14), (8, 16), (3, 10)])
def test_mg16_prove_one_call(be, ndev, log_n):
    if gpu_count() < ndev:
        pytest.skip("needs %d GPUs" % ndev)
    tl = P.multi_groth16_case(L.load(), be, ndev, log_n, seed=log_n + ndev, devices=list(range(ndev)))
    assert ndev == 1 or len(tl) >= 3


def test_mg16_prove_dense(be):
    P.multi_groth16_dense_case(L.load(), be, min(2, gpu_count()), devices=list(range(min(2, gpu_count()))))


@pytest.mark.parametrize("ndev,group,n", [(1, L.PS_G1, 3000), (2, L.PS_G1, 1 << 16), (2, L.PS_G2, 5000), (8, L.PS_G1, 1 << 18)])
def test_mmsm_one_call(ndev, group, n):
    if gpu_count() < ndev:
        pytest.skip("needs %d GPUs" % ndev)
    P.multi_msm_case(L.load(), ndev, group, n, devices=list(range(ndev)), tables=-1)


@pytest.mark.parametrize("world", [2, 4])
def test_dist_nccl(world):
    """one process per GPU over NCCL (torchrun): groth16_prove_sharded and msm_sharded against the exponent-level
    expectation, including the pipelined flow"""
    if gpu_count() < world:
        pytest.skip("needs %d GPUs" % world)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + world), os.path.join(ROOT, "tests", "nccl_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "NCCL_WORKER_OK" in res.stdout
