#!/usr/bin/env python3
"""Window model check: MSM time against the bucket cost used by the automatic window choice.
Usage: window_sweep.py [costs] [g1_logs] [g2_logs]"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import playsnark_b200 as ps  # noqa: E402
from playsnark_b200 import _lib as L  # noqa: E402

costs = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "38,100,160").split(",")]
g1 = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "16,18,20,22").split(",") if x]
g2 = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "16,18,20").split(",") if x]
be = ps.Backend(0)
rng = np.random.default_rng(5)


def rand_scalars(n):
    a = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    a[:, 0] &= 0x3F
    return a.tobytes()


rows = []
for group, name, logs in ((L.PS_G1, "G1", g1), (L.PS_G2, "G2", g2)):
    for log_n in logs:
        n = 1 << log_n
        ks, sc = rand_scalars(n), rand_scalars(n)
        seen = {}
        for cost in costs:
            be.set_option("msm_bucket_cost", cost)
            bases = be.bases_from_scalars(group, ks, 0, -1)
            info = (ctypes.c_int * 4)()
            be.lib.ps_bases_info(bases.handle, info)
            if info[0] not in seen:
                res = be.msm(bases, sc)
                best = None
                for _ in range(3):
                    be.msm(bases, sc)
                    t = be.msm_timing()
                    if best is None or t["total_ms"] < best["total_ms"]:
                        best = t
                seen[info[0]] = (res, best)
            res, best = seen[info[0]]
            row = dict(group=name, log_n=log_n, bucket_cost=cost, c=info[0], W=info[1], **{k: round(v, 3) for k, v in best.items()})
            rows.append(row)
            print(row, flush=True)
            bases.close()
        assert len({r for r, _ in seen.values()}) == 1, "results differ between windows"
be.set_option("msm_bucket_cost", 100)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "window_sweep.json"), "w"), indent=1)
