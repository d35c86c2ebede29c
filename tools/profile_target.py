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
    import numpy as np
    g = np.random.default_rng(1)

    def scalars(seed_off):       # uniform below 2^254 < r, big-endian rows (as bench.py)
        a = g.integers(0, 256, size=(n, 32), dtype=np.uint8)
        a[:, 0] &= 0x3F
        a[:, 31] |= 1
        return a.tobytes()
    ks, sc = scalars(0), scalars(1)
    bases = be.bases_from_scalars(group, ks, 0, -1)   # all window tables, like the keys and bench.py
    for _ in range(2):
        be.msm(bases, sc)
    print(be.msm_timing())
if what == "mul":
    v, ms = C.c_double(), C.c_double()
    be._check(be.lib.ps_bench_fieldmul(be.ctx, 1, 500, C.byref(v), C.byref(ms)))
    print("fp mul/s %.3e" % v.value)
