"""torchrun worker of tests/test_gpu_multi.py::test_dist_nccl: every rank owns one GPU; rank 0 checks the results."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist

from oracle import ps_oracle as O
from playsnark_b200 import _lib as L, api, dist as D
from tests import helpers as H


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    be = api.Backend(local)
    be.set_stream(torch.cuda.current_stream().cuda_stream)
    # sharded Groth16 (pipelined flow for world = 2 * 2^j)
    n = 1 << 12
    sq, wit = H.sparse_circuit(n, 5, n // 2)
    tr, tw = H.sparse_groth16_setup(be, sq, 5)
    smp = O.Sampler(77)
    r, s = smp.fr(), smp.fr()
    for split in (True, False):
        got = D.groth16_prove_sharded(be, tr, sq, wit, r, s, dist, dev, split_quotient=split)
        torch.cuda.synchronize()
        if rank == 0:
            A, B, Cc, _ = H.sparse_groth16_expected(sq, wit, tw, r, s)
            assert tuple(got) == (A, B, Cc), "sharded proof differs (split=%s)" % split
    # sharded MSM: every rank its own point range
    import random
    rng = random.Random(1000 + rank)
    m = 5000
    ks = [rng.randrange(1, O.R) for _ in range(m)]
    sc = [rng.randrange(O.R) for _ in range(m)]
    bases = be.bases_from_scalars(L.PS_G1, ks)
    import numpy as np
    le = np.frombuffer(b"".join(v.to_bytes(32, "little") for v in sc), dtype=np.int32).reshape(m, 8).copy()
    d_sc = torch.from_numpy(le).to(dev)
    out = D.msm_sharded(be, bases, d_sc, m, dist)
    exps = [None] * world
    dist.all_gather_object(exps, sum(k * v for k, v in zip(ks, sc)) % O.R)
    if rank == 0:
        assert out == O.g1_compress(O.g1_mul(sum(exps) % O.R)), "sharded MSM differs"
        print("NCCL_WORKER_OK world=%d" % world, flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
