#!/usr/bin/env python3
"""A/B of the merged Y3 of the group law (curve.cuh mul_sub_pair: a b - c d as ONE reduction of two unreduced products)
inside the bucket-accumulation kernels on one B200, against the product library:
  g1nolazyy3   G1: Y3 as two Montgomery products (group_g1.cu with -DPS_NO_LAZY_Y3)
G2 MSMs at 2^18 / 2^20 points, G1 at 2^24, all window tables, every result checked against (sum k_i s_i mod r) * G from
the oracle.  Each library runs in its own process (PLAYSNARK_B200_LIB).  Round 2's other variants (unreduced Fp2
products, Fp2 Y3 on three differences, wide squaring) were measured with this tool at commit a58c4da and removed;
numbers in profiles/r02_ab_lazy.md.
    python tools/ab_lazy.py --build    # here: the variants into playsnark_b200/variants/
    python tools/ab_lazy.py            # on the GPU box
"""
import ctypes as C
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import numpy as np
    import torch
    import playsnark_b200 as ps
    from playsnark_b200 import _lib as L
    import bench as B
    be = ps.Backend(0)
    lib = be.lib
    dev = torch.device("cuda:0")
    res = {}
    cases = [(L.PS_G2, 18), (L.PS_G2, 20), (L.PS_G1, 24)]
    only = os.environ.get("AB_ONLY")
    if only:
        cases = [c for c in cases if c[0] == int(only)]
    for group, k in cases:
        n = 1 << k
        ks, sc = B.random_scalars_be(n, 3000 + 16 * k), B.random_scalars_be(n, 4000 + 16 * k)
        bases = be.bases_from_scalars(group, ks.tobytes(), 0, -1)
        d_sc = torch.from_numpy(B.be_to_le_limbs(sc).view(np.int32)).to(dev)
        d_part = torch.zeros(384, dtype=torch.uint8, device=dev)
        step = lambda: be._check(lib.ps_msm_device(be.ctx, bases.handle, 0, C.c_void_p(d_sc.data_ptr()), n, C.c_void_p(d_part.data_ptr())))
        for _ in range(3):
            step()
        be.sync()
        out = C.create_string_buffer(96 if group == L.PS_G2 else 48)
        be._check(lib.ps_msm_combine(be.ctx, group, C.c_void_p(d_part.data_ptr()), 1, out))
        ok = out.raw == B.expected_point(group, B.expected_exponent(ks, sc))
        best, acc = 1e9, 1e9
        for _ in range(7):
            step()
            be.sync()
            t = be.msm_timing()
            best, acc = min(best, t["total_ms"]), min(acc, t["accumulate_ms"])
        # rare-event check: more scalar vectors over the same bases, each against the oracle's expectation
        extra = 0
        for rnd in range(int(os.environ.get("AB_EXTRA_PARITY", "4")) if k >= 22 else 0):
            sc_x = B.random_scalars_be(n, 9000 + 7 * rnd + k)
            d_sc.copy_(torch.from_numpy(B.be_to_le_limbs(sc_x).view(np.int32)).to(dev))
            step()
            be._check(lib.ps_msm_combine(be.ctx, group, C.c_void_p(d_part.data_ptr()), 1, out))
            ok = ok and out.raw == B.expected_point(group, B.expected_exponent(ks, sc_x))
            extra += 1
        res["g%d_2p%d" % (group, k)] = {"extra_parity_rounds": extra, "total_ms": best, "accumulate_ms": acc, "parity": ok, "point": out.raw.hex()[:16]}
        bases.close()
        del d_sc
        torch.cuda.empty_cache()
    print("AB_RESULT " + json.dumps(res), flush=True)


def main():
    if "--child" in sys.argv:
        return child()
    if "--build" in sys.argv:
        from playsnark_b200 import build as B
        B.build_variant("g1nolazyy3", ["-DPS_NO_LAZY_Y3"], tus=("group_g1.cu",))
        return
    libs = [("product", os.path.join(ROOT, "playsnark_b200", "libplaysnark_b200.so"))]
    for p in sorted(glob.glob(os.path.join(ROOT, "playsnark_b200", "variants", "lib_fp2*.so")) +
                    glob.glob(os.path.join(ROOT, "playsnark_b200", "variants", "lib_g1*.so"))):   # fp2*: G2-only variants
        libs.append((os.path.basename(p)[4:-3], p))
    rows = {}
    for name, path in libs:
        env = dict(os.environ, PLAYSNARK_B200_LIB=path)
        if name.startswith("fp2"):
            env["AB_ONLY"] = "2"     # the variant only changes the G2 kernel
        elif name.startswith("g1"):
            env["AB_ONLY"] = "1"
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=env, capture_output=True, text=True, timeout=600)
        line = [l for l in out.stdout.splitlines() if l.startswith("AB_RESULT ")]
        rows[name] = json.loads(line[0][10:]) if line else {"error": out.stderr[-800:]}
        print(name, json.dumps(rows[name]), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "ab_lazy.json"), "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
