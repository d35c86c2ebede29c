#!/usr/bin/env python3
"""Development probe (run under gpurun): integer-pipe peaks, field-mul rates, MSM phase timings."""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import playsnark_b200 as ps  # noqa: E402
from playsnark_b200 import _lib as L  # noqa: E402


def main():
    args = sys.argv[1:]
    max_log = int(args[0]) if args else 20
    be = ps.Backend(0)
    lib = be.lib
    out = {}
    v, ms = C.c_double(), C.c_double()
    for variant, name in enumerate(["imad_lo", "imad_hi", "imad_wide", "imad_wide_carry"]):
        be._check(lib.ps_bench_intpipe(be.ctx, variant, 2000, C.byref(v), C.byref(ms)))
        out[name] = {"inst_per_s": v.value, "ms": ms.value}
        print("intpipe %-16s %.3e inst/s  (%.2f ms)" % (name, v.value, ms.value), flush=True)
    for field, name in enumerate(["fr", "fp"]):
        be._check(lib.ps_bench_fieldmul(be.ctx, field, 2000, C.byref(v), C.byref(ms)))
        out["mul_" + name] = {"mul_per_s": v.value, "ms": ms.value}
        print("fieldmul %-3s %.3e mul/s (%.2f ms)" % (name, v.value, ms.value), flush=True)
    import random
    rng = random.Random(1)
    R = ps.R
    for group, gname, top in ((L.PS_G1, "g1", max_log), (L.PS_G2, "g2", max(12, max_log - 3))):
        for log_n in range(12, top + 1, 2):
            n = 1 << log_n
            ks = b"".join(rng.randrange(1, R).to_bytes(32, "big") for _ in range(n))
            sc = b"".join(rng.randrange(R).to_bytes(32, "big") for _ in range(n))
            for cfg, wb, tables in (("T1", 0, 1), ("full", 0, -1)):
                t0 = time.time()
                bases = be.bases_from_scalars(group, ks, wb, tables)
                be.sync()
                t_bases = time.time() - t0
                be.msm(bases, sc)  # warm-up
                t0 = time.time()
                be.msm(bases, sc)
                wall = time.time() - t0
                tm = be.msm_timing()
                out["msm_%s_%d_%s" % (gname, log_n, cfg)] = dict(tm, wall_ms=wall * 1e3, bases_s=t_bases)
                print("msm %s 2^%d %-4s: total %.3f ms (sort %.3f, accum %.3f, comb %.3f, reduce %.3f) wall %.1f ms; %.3e pts/s; bases %.2fs" % (
                    gname, log_n, cfg, tm["total_ms"], tm["sort_ms"], tm["accumulate_ms"], tm["combine_ms"], tm["reduce_ms"], wall * 1e3,
                    n / (tm["total_ms"] * 1e-3), t_bases), flush=True)
                bases.close()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
