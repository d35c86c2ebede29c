#!/usr/bin/env python3
"""A/B of the Fp2 product inside the G2 bucket-accumulation kernel on one B200: Karatsuba with three Montgomery products
(product library) against Karatsuba on unreduced products with two reductions (variants/lib_fp2lazy.so, accum_g2.cu built
with -DPS_FP2_LAZY).  G2 MSMs at 2^18 and 2^20 points with all window tables, every result checked against
(sum k_i s_i mod r) * G2 from the oracle.  Each library runs in its own process (PLAYSNARK_B200_LIB).
    python -c "from playsnark_b200 import build as B; B.build_variant('fp2lazy', ['-DPS_FP2_LAZY'], tus=('accum_g2.cu',))"   # here
    python tools/ab_fp2.py            # on the GPU box
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
    for k in (18, 20):
        n = 1 << k
        ks, sc = B.random_scalars_be(n, 3000 + 16 * k), B.random_scalars_be(n, 4000 + 16 * k)
        bases = be.bases_from_scalars(L.PS_G2, ks.tobytes(), 0, -1)
        d_sc = torch.from_numpy(B.be_to_le_limbs(sc).view(np.int32)).to(dev)
        d_part = torch.zeros(384, dtype=torch.uint8, device=dev)
        step = lambda: be._check(lib.ps_msm_device(be.ctx, bases.handle, 0, C.c_void_p(d_sc.data_ptr()), n, C.c_void_p(d_part.data_ptr())))
        for _ in range(3):
            step()
        be.sync()
        out = C.create_string_buffer(96)
        be._check(lib.ps_msm_combine(be.ctx, L.PS_G2, C.c_void_p(d_part.data_ptr()), 1, out))
        ok = out.raw == B.expected_point(L.PS_G2, B.expected_exponent(ks, sc))
        best, acc = 1e9, 1e9
        for _ in range(7):
            step()
            be.sync()
            t = be.msm_timing()
            best, acc = min(best, t["total_ms"]), min(acc, t["accumulate_ms"])
        res["2p%d" % k] = {"total_ms": best, "accumulate_ms": acc, "parity": ok, "point": out.raw.hex()[:16]}
        bases.close()
    print("AB_RESULT " + json.dumps(res), flush=True)


def main():
    if "--child" in sys.argv:
        return child()
    libs = [("product", os.path.join(ROOT, "playsnark_b200", "libplaysnark_b200.so"))]
    for p in sorted(glob.glob(os.path.join(ROOT, "playsnark_b200", "variants", "lib_fp2*.so"))):
        libs.append((os.path.basename(p)[4:-3], p))
    rows = {}
    for name, path in libs:
        env = dict(os.environ, PLAYSNARK_B200_LIB=path)
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=env, capture_output=True, text=True, timeout=600)
        line = [l for l in out.stdout.splitlines() if l.startswith("AB_RESULT ")]
        rows[name] = json.loads(line[0][10:]) if line else {"error": out.stderr[-800:]}
        print(name, json.dumps(rows[name]), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "ab_fp2.json"), "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
