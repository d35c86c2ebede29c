#!/usr/bin/env python3
"""XYZZ-chain vs batched-affine bucket accumulation (development probe)."""
import os, sys, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import playsnark_b200 as ps
from playsnark_b200 import _lib as L
be = ps.Backend(0)
rng = np.random.default_rng(5)
def rs(n):
    a = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); a[:, 0] &= 0x3F; return a.tobytes()
cases = [(L.PS_G1, int(x)) for x in sys.argv[1].split(",")] + [(L.PS_G2, int(x)) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else [])]
for grp, logn in cases:
    n = 1 << logn
    b = be.bases_from_scalars(grp, rs(n), 0, -1)
    sc = rs(n)
    outs = []
    for mode in (0, 1):
        be.set_option("msm_accumulate", mode)
        r = be.msm(b, sc); r = be.msm(b, sc)
        outs.append(r)
        print("group %d 2^%d mode %d:" % (grp, logn, mode), {k: round(v, 2) for k, v in be.msm_timing().items()}, flush=True)
    assert outs[0] == outs[1], "results differ between accumulation modes"
    b.close()
print("same results in both modes")
