#!/usr/bin/env python3
"""Sharded Groth16 under torchrun: proof latency for several MSM shares of rank 0 (which also divides),
with parity against the exponent-level expectation.
  torchrun --nproc-per-node N tools/g16_sharded_probe.py <log_n> [shares, comma separated]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import playsnark_b200 as ps  # noqa: E402
from playsnark_b200 import dist as D  # noqa: E402
from oracle import ps_oracle as O  # noqa: E402
from tests import helpers as H  # noqa: E402

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
shares = [float(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "0.3,0.5,0.75,1.0").split(",")]
local = int(os.environ.get("LOCAL_RANK", 0))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
rank, world = dist.get_rank(), dist.get_world_size()
be = ps.Backend(local)
be.set_stream(torch.cuda.current_stream().cuda_stream)
n = 1 << log_n
sq, wit = H.sparse_circuit(n, 7, n // 2)
tr, tw = H.sparse_groth16_setup(be, sq, 7)
smp = O.Sampler(99)
r, s = smp.fr(), smp.fr()
wb = ps.HostBuffer(be, b"".join(v.to_bytes(32, "big") for v in wit))
sq._resident(be); D.load_key_sharded(be, tr, world); be.sync()
want = H.sparse_groth16_expected(sq, wit, tw, r, s)[:3] if rank == 0 else None
for share in shares:
    for _ in range(2):
        pr = D.groth16_prove_sharded(be, tr, sq, wb, r, s, dist, dev, rank0_share=share)
    dist.barrier(); torch.cuda.synchronize()
    reps = 5
    t0 = time.perf_counter()
    for _ in range(reps):
        pr = D.groth16_prove_sharded(be, tr, sq, wb, r, s, dist, dev, rank0_share=share)
    dist.barrier(); torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / reps * 1e3
    if rank == 0:
        print("world %d 2^%d rank0_share %.2f: %.2f ms per proof, parity %s" % (world, log_n, share, ms, tuple(pr) == tuple(want)), flush=True)
# stage timeline of one proof per rank (CUDA events; ms since the start of the call)
trace = []
pr = D.groth16_prove_sharded(be, tr, sq, wb, r, s, dist, dev, rank0_share=shares[0], trace=trace)
torch.cuda.synchronize()
line = "rank %d: " % rank + ", ".join("%s %.2f" % (nm, trace[0][1].elapsed_time(ev)) for nm, ev in trace[1:])
for rk in range(world):
    dist.barrier()
    if rk == rank:
        print(line, flush=True)
dist.destroy_process_group()
