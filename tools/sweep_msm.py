#!/usr/bin/env python3
"""MSM sweep for profiles/: G1 and G2, 2^12..2^max, random and witness-like scalar distributions.
Usage: sweep_msm.py <g1_max_log> <g2_max_log>"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import playsnark_b200 as ps  # noqa: E402
from playsnark_b200 import _lib as L  # noqa: E402

g1_max = int(sys.argv[1]) if len(sys.argv) > 1 else 22
g2_max = int(sys.argv[2]) if len(sys.argv) > 2 else 20
be = ps.Backend(0)
rng = np.random.default_rng(5)


def rand_scalars(n):
    a = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    a[:, 0] &= 0x3F
    return a.tobytes()


def kind_scalars(kind, n):
    if kind == "rand":
        return rand_scalars(n)
    if kind == "ones":
        return (1).to_bytes(32, "big") * n
    if kind == "minus_one":
        return (ps.R - 1).to_bytes(32, "big") * n
    if kind == "bits":
        a = np.zeros((n, 32), dtype=np.uint8)
        a[:, 31] = rng.integers(0, 2, size=n, dtype=np.uint8)
        return a.tobytes()
    if kind == "small64":
        a = np.zeros((n, 32), dtype=np.uint8)
        a[:, 24:] = rng.integers(0, 256, size=(n, 8), dtype=np.uint8)
        return a.tobytes()
    raise ValueError(kind)


rows = []
for group, name, top in ((L.PS_G1, "G1", g1_max), (L.PS_G2, "G2", g2_max)):
    for log_n in range(12, top + 1, 2):
        n = 1 << log_n
        bases = be.bases_from_scalars(group, rand_scalars(n), 0, -1)
        info = (__import__("ctypes").c_int * 4)()
        be.lib.ps_bases_info(bases.handle, info)
        kinds = ["rand"] + (["ones", "minus_one", "bits", "small64"] if log_n in (16, 20) else [])
        for kind in kinds:
            sc = kind_scalars(kind, n)
            be.msm(bases, sc)
            best = None
            for _ in range(3):
                be.msm(bases, sc)
                t = be.msm_timing()
                if best is None or t["total_ms"] < best["total_ms"]:
                    best = t
            row = dict(group=name, log_n=log_n, scalars=kind, c=info[0], W=info[1], T=info[2], **{k: round(v, 3) for k, v in best.items()})
            row["points_per_s"] = n / (best["total_ms"] * 1e-3)
            rows.append(row)
            print(row, flush=True)
        bases.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "msm_sweep.json"), "w"), indent=1)
