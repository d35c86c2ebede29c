#!/usr/bin/env python3
"""A/B of the counting sort's scatter: one pass vs two passes through a partitioned staging array.
Usage: ab_scatter.py [g1_logs]"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import playsnark_b200 as ps  # noqa: E402
from playsnark_b200 import _lib as L  # noqa: E402

logs = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "22,24").split(",") if x]
be = ps.Backend(0)
rng = np.random.default_rng(5)


def rand_scalars(n):
    a = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    a[:, 0] &= 0x3F
    return a.tobytes()


rows = []
for log_n in logs:
    n = 1 << log_n
    bases = be.bases_from_scalars(L.PS_G1, rand_scalars(n), 0, -1)
    info = (ctypes.c_int * 4)()
    be.lib.ps_bases_info(bases.handle, info)
    sc = rand_scalars(n)
    outs = {}
    for mode in (0, 2):
        be.set_option("msm_scatter", mode)
        outs[mode] = be.msm(bases, sc)
        best = None
        for _ in range(3):
            be.msm(bases, sc)
            t = be.msm_timing()
            if best is None or t["total_ms"] < best["total_ms"]:
                best = t
        row = dict(group="G1", log_n=log_n, scatter=mode, c=info[0], W=info[1], **{k: round(v, 3) for k, v in best.items()})
        rows.append(row)
        print(row, flush=True)
    assert outs[0] == outs[2], "one-pass and two-pass scatter disagree"
    bases.close()
be.set_option("msm_scatter", 1)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "ab_scatter.json"), "w"), indent=1)
