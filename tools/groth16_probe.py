#!/usr/bin/env python3
"""Development probe: Groth16 prove on a sparse synthetic circuit of 2^k gates, with timing and the
exponent-level parity check.  Usage: groth16_probe.py <log_n> [reps]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import playsnark_b200 as ps  # noqa: E402
from oracle import ps_oracle as O  # noqa: E402
from tests import helpers as H  # noqa: E402

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n = 1 << log_n
be = ps.Backend(0)
t0 = time.time(); sq, wit = H.sparse_circuit(n, 7, n // 2); print("circuit %.1fs" % (time.time() - t0), flush=True)
t0 = time.time(); tr, tw = H.sparse_groth16_setup(be, sq, 7); print("setup %.1fs" % (time.time() - t0), flush=True)
smp = O.Sampler(99); r, s = smp.fr(), smp.fr()
t0 = time.time(); sq._resident(be); be.sync(); print("qap load %.2fs" % (time.time() - t0), flush=True)
t0 = time.time(); tr._resident(be); be.sync(); print("key load %.2fs" % (time.time() - t0), flush=True)
wb = b"".join(v.to_bytes(32, "big") for v in wit)
for i in range(reps):
    l0 = be.launch_count()
    t0 = time.time(); pr = ps.Groth16Prove(tr, sq, wb, r, s, backend=be); dt = time.time() - t0
    print("prove %d: %.2f ms wall (witness pre-marshalled), %d launches, device phases %s" % (
        i, dt * 1e3, be.launch_count() - l0, {k: round(v, 2) for k, v in be.prove_timing().items()}), flush=True)
hb = ps.HostBuffer(be, wb)
for i in range(reps):
    t0 = time.time(); pr_p = ps.Groth16Prove(tr, sq, hb, r, s, backend=be); dt = time.time() - t0
    print("prove (pinned witness) %d: %.2f ms wall, device phases %s" % (i, dt * 1e3, {k: round(v, 2) for k, v in be.prove_timing().items()}), flush=True)
assert (pr_p.A, pr_p.B, pr_p.C) == (pr.A, pr.B, pr.C)
print("last MSM inside prove (B, G2):", {k: round(v, 2) for k, v in be.msm_timing().items()})
import random as _r
from playsnark_b200 import _lib as _L
_rng = _r.Random(5)
_sc = b"".join(_rng.randrange(ps.R).to_bytes(32, "big") for _ in range(n + 2))
for grp in (_L.PS_G1, _L.PS_G2):
    _b = be.bases_from_scalars(grp, _sc, 0, -1)
    be.msm(_b, _sc); be.msm(_b, _sc)
    info = (__import__("ctypes").c_int * 4)(); be.lib.ps_bases_info(_b.handle, info)
    print("standalone group %d n+2 points c=%d W=%d T=%d:" % (grp, info[0], info[1], info[2]), {k: round(v, 2) for k, v in be.msm_timing().items()})
t0 = time.time(); h, _ = ps.Quotient(sq, wit, backend=be, return_abc=True); print("quotient alone %.2f ms" % ((time.time() - t0) * 1e3))
A, B, C, _ = H.sparse_groth16_expected(sq, wit, tw, r, s)
assert (pr.A, pr.B, pr.C) == (A, B, C), "proof differs from the exponent-level expectation"
print("parity ok at 2^%d" % log_n)
