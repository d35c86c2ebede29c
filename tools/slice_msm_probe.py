#!/usr/bin/env python3
"""One-GPU proxy for the per-device MSM phases of the multi-GPU Groth16 prover (ps_mg16_prove): the key of a 2^k-constraint
circuit is loaded with windows sized for a 1/world share (option msm_shards) and ps_g16_msm_partials is timed on the
index ranges one device of `world` would own -- early phase (A_d, B_d on the second stream, the tail of C_d) and late
phase (h_d . XiT_d) -- plus each group alone, so that the cost of running G1 and G2 side by side is visible.
  python tools/slice_msm_probe.py [log_n=20] [world=8]"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import playsnark_b200 as ps  # noqa: E402
from playsnark_b200 import synth, dist as D, api  # noqa: E402
from oracle import ps_oracle as O  # noqa: E402

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
n = 1 << log_n
dev = torch.device("cuda", 0)
be = ps.Backend(0)
be.set_stream(torch.cuda.current_stream().cuda_stream)
lib = be.lib
sq, wit = synth.sparse_circuit(n, 7, n // 2)
smp = O.Sampler(99)
toxic = tuple(smp.fr() for _ in range(5))
r, s = smp.fr(), smp.fr()
be.set_option("msm_shards", world)
tr = ps.NewGroth16TrustedSetup(sq, backend=be, toxic=toxic, export=False)
be.set_option("msm_shards", 1)
kh, qh = tr._resident(be), sq._resident(be)
nA, nC, nB = (int(lib.ps_g16_scalar_count(kh, w)) for w in (0, 1, 2))
nio = sq.nbIO
head = nio + n - 1
new = lambda rows: torch.zeros((rows, 8), dtype=torch.int32, device=dev)
ptr = lambda t: C.c_void_p(t.data_ptr())
bufA, bufC, bufB = new(nA), new(nC), new(nB)
wb = api._fr_bytes(wit)
be._check(lib.ps_g16_scalars(be.ctx, kh, qh, wb, api._fr_bytes([r]), api._fr_bytes([s]), ptr(bufA), ptr(bufC), ptr(bufB)))
be.sync()
weights = [1.0] * world
k = world - 1   # the last device's ranges
rA = D.weighted_ranges(nA, weights)[k]
rB = D.weighted_ranges(nB, weights)[k]
rT = D.weighted_ranges(nC - head, weights)[k]
rH = D.weighted_ranges(head, weights)[k]
out = torch.zeros(768, dtype=torch.uint8, device=dev)


def partials(sp):
    first = (C.c_size_t * 3)(*[lo for lo, _ in sp])
    cnt = (C.c_size_t * 3)(*[hi - lo for lo, hi in sp])
    views = [buf[lo:hi] if hi > lo else buf for buf, (lo, hi) in zip((bufA, bufC, bufB), sp)]
    be._check(lib.ps_g16_msm_partials(be.ctx, kh, ptr(views[0]), ptr(views[1]), ptr(views[2]), first, cnt, ptr(out)))


def timed(sp, reps=10):
    for _ in range(3):
        partials(sp)
    be.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        partials(sp)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


Z = (0, 0)
tail = (head + rT[0], head + rT[1])
res = {"log_n": log_n, "world": world, "points": {"A": rA[1] - rA[0], "C_tail": rT[1] - rT[0], "B": rB[1] - rB[0], "h": rH[1] - rH[0]}}
for wf in (0, 1):       # accumulate chunks: wave count rounded up (old) / down (default)
    be.set_option("msm_wave_floor", wf)
    row = {}
    row["early_all_ms"] = timed([rA, tail, rB])
    row["early_g1_only_ms"] = timed([rA, tail, Z])
    row["early_g2_only_ms"] = timed([Z, Z, rB])
    row["A_only_ms"] = timed([rA, Z, Z])
    row["late_h_ms"] = timed([Z, rH, Z])
    row["msm_timing_last"] = be.msm_timing()
    res["wave_floor_%d" % wf] = row
info = (C.c_int * 4)()
for nm, w in (("A", 0), ("C", 1), ("B", 2)):
    lib.ps_bases_info(lib.ps_g16_key_bases(kh, w), info)
    res["window_" + nm] = {"c": info[0], "W": info[1]}
print(json.dumps(res, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "slice_msm_probe_%d_%d.json" % (log_n, world)), "w") as f:
    json.dump(res, f, indent=1)
