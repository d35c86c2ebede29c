#!/usr/bin/env python3
"""One-call multi-GPU Groth16 (ps_mg16_prove) on all visible GPUs of the box: proof latency and stage timeline for several
MSM shares of device 0 (which also divides), key made by the device setup, exponent-level parity on every configuration.
  python tools/mg16_probe.py [log_n=20] [shares in percent, comma separated; 0 = the library's default]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import playsnark_b200 as ps  # noqa: E402
from playsnark_b200 import synth  # noqa: E402
from oracle import expect as E, ps_oracle as O  # noqa: E402

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
shares = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "0,40,55,70,85").split(",")]
world = torch.cuda.device_count()
n = 1 << log_n
be = ps.Backend(0)
sq, wit = synth.sparse_circuit(n, 7, n // 2)
smp = O.Sampler(99)
toxic = tuple(smp.fr() for _ in range(5))
r, s = smp.fr(), smp.fr()
tr = ps.NewGroth16TrustedSetup(sq, backend=be, toxic=toxic, export=True)
sq.close(); tr.close()
want = E.groth16_expected(sq, wit, toxic, r, s)[:3]
mb = ps.MultiBackend(list(range(world)))
wb = ps.HostBuffer(be, b"".join(v.to_bytes(32, "big") for v in wit))
marks = ["witness gathered", "subtree interpolated", "roots gathered", "top levels", "a, b swapped", "slice scalars (+ division on device 0)",
         "MSMs (B_d at once, A_d + C_d after h)", "record ready", "combined + encoded"]
rows = []
sq._resident(mb)
for share, wave_floor in [(shares[0], 0)] + [(sh, 1) for sh in shares]:
    tr.close()
    mb.set_option("rank0_share_percent", share)
    mb.set_option("msm_wave_floor", wave_floor)
    tr._resident(mb)
    for _ in range(3):
        pr = ps.Groth16Prove(tr, sq, wb, r, s, backend=mb)
    reps, best, t_all = 8, 1e9, time.perf_counter()
    for _ in range(reps):
        t0 = time.perf_counter()
        pr = ps.Groth16Prove(tr, sq, wb, r, s, backend=mb)
        best = min(best, time.perf_counter() - t0)
    avg = (time.perf_counter() - t_all) / reps
    row = {"world": world, "log_n": log_n, "rank0_share_percent": share, "msm_wave_floor": wave_floor, "ms_avg": avg * 1e3, "ms_best": best * 1e3,
           "parity": (pr.A, pr.B, pr.C) == tuple(want),
           "timeline_ms": {"device%d" % d: dict(zip(marks, [round(x, 3) for x in mb.timeline(d)])) for d in sorted({0, 1, world - 1})}}
    rows.append(row)
    print(json.dumps(row), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "mg16_probe_%dgpu_2p%d.json" % (world, log_n)), "w") as f:
    json.dump(rows, f, indent=1)
wb.close(); sq.close(); tr.close(); mb.close()
