#!/usr/bin/env python3
"""Short Groth16 workload for ncu captures: sparse circuit of 2^k constraints, 2 proofs."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import playsnark_b200 as ps  # noqa: E402
from oracle import ps_oracle as O  # noqa: E402
from playsnark_b200 import synth  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else 18
n = 1 << k
be = ps.Backend(0)
sq, wit = synth.sparse_circuit(n, 7, n // 2)
smp0 = O.Sampler(7)
tr = ps.NewGroth16TrustedSetup(sq, backend=be, toxic=tuple(smp0.fr() for _ in range(5)), export=False)   # key made on the device
wb = ps.HostBuffer(be, b"".join(v.to_bytes(32, "big") for v in wit))
smp = O.Sampler(99)
r, s = smp.fr(), smp.fr()
sq._resident(be); tr._resident(be); be.sync()
print("LOADED launches=%d" % be.launch_count(), flush=True)
for i in range(2):
    l0 = be.launch_count()
    ps.Groth16Prove(tr, sq, wb, r, s, backend=be)
    print("proof %d launches %d..%d" % (i, l0, be.launch_count()), be.prove_timing(), flush=True)
