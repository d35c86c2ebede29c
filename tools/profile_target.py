#!/usr/bin/env python3
"""Short workload for ncu captures: one G1 MSM (and optionally G2 / the field-mul microbench)."""
import ctypes as C
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import playsnark_b200 as ps  # noqa: E402
from playsnark_b200 import _lib as L  # noqa: E402

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
what = sys.argv[2] if len(sys.argv) > 2 else "g1"
be = ps.Backend(0)
rng = random.Random(1)
n = 1 << log_n
if what in ("g1", "g2"):
    group = L.PS_G1 if what == "g1" else L.PS_G2
    ks = b"".join(rng.randrange(1, ps.R).to_bytes(32, "big") for _ in range(n))
    sc = b"".join(rng.randrange(ps.R).to_bytes(32, "big") for _ in range(n))
    bases = be.bases_from_scalars(group, ks, 0, -1)   # all window tables, like the keys and bench.py
    for _ in range(2):
        be.msm(bases, sc)
    print(be.msm_timing())
if what == "mul":
    v, ms = C.c_double(), C.c_double()
    be._check(be.lib.ps_bench_fieldmul(be.ctx, 1, 500, C.byref(v), C.byref(ms)))
    print("fp mul/s %.3e" % v.value)
