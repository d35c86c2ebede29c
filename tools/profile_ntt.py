#!/usr/bin/env python3
"""Short NTT workload for ncu captures: forward + inverse transform of 2^k Fr elements through ps_ntt_fr."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C
import numpy as np
import playsnark_b200 as ps
k = int(sys.argv[1]) if len(sys.argv) > 1 else 22
n = 1 << k
be = ps.Backend(0)
rng = np.random.default_rng(1)
a = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); a[:, 0] &= 0x3F
buf = C.create_string_buffer(a.tobytes(), n * 32)
for inv in (0, 1, 0, 1):
    be._check(be.lib.ps_ntt_fr(be.ctx, buf, k, inv, None))
assert buf.raw == a.tobytes(), "NTT round trip failed"
print("ntt round trip ok, launches", be.launch_count())
