#!/usr/bin/env python3
"""A/B of NTT kernel variants on one B200: quotient time (ps_g16_prove's first phase) at 2^16 and 2^20 constraints and
raw NTT passes, for every library under playsnark_b200/variants/ (built here with playsnark_b200.build.build_variant)
and the product library.  Each variant runs in its own process (PLAYSNARK_B200_LIB).  Usage on the GPU box:
    python tools/ab_ntt.py            # all variants
    python tools/ab_ntt.py --child    # (internal) one measurement in this process
"""
import glob
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import playsnark_b200 as ps
    from playsnark_b200 import synth
    from oracle import ps_oracle as O
    be = ps.Backend(0)
    res = {}
    for k in (16, 20):
        n = 1 << k
        sq, wit = synth.sparse_circuit(n, 7, n // 2)
        smp = O.Sampler(99)
        toxic = tuple(smp.fr() for _ in range(5))
        r, s = smp.fr(), smp.fr()
        tr = ps.NewGroth16TrustedSetup(sq, backend=be, toxic=toxic, export=False)
        wb = ps.HostBuffer(be, b"".join(v.to_bytes(32, "big") for v in wit))
        first = None
        q, tot, wall = [], [], []
        for i in range(7):
            t0 = time.perf_counter()
            pr = ps.Groth16Prove(tr, sq, wb, r, s, backend=be)
            wall.append((time.perf_counter() - t0) * 1e3)
            t = be.prove_timing()
            q.append(t["quotient_ms"]); tot.append(t["total_ms"])
            first = first or (pr.A, pr.B, pr.C)
            assert (pr.A, pr.B, pr.C) == first
        res["2p%d" % k] = {"quotient_ms": min(q), "device_total_ms": min(tot), "wall_ms": min(wall), "proof": first[0].hex()[:16]}
        wb.close(); tr.close(); sq.close()
    print("AB_RESULT " + json.dumps(res), flush=True)


def main():
    if "--child" in sys.argv:
        return child()
    libs = [("product", os.path.join(ROOT, "playsnark_b200", "libplaysnark_b200.so"))]
    for p in sorted(glob.glob(os.path.join(ROOT, "playsnark_b200", "variants", "lib_*.so"))):
        libs.append((os.path.basename(p)[4:-3], p))
    rows = {}
    for name, path in libs:
        env = dict(os.environ, PLAYSNARK_B200_LIB=path)
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=env, capture_output=True, text=True, timeout=900)
        line = [l for l in out.stdout.splitlines() if l.startswith("AB_RESULT ")]
        rows[name] = json.loads(line[0][10:]) if line else {"error": out.stderr[-500:]}
        print(name, json.dumps(rows[name]), flush=True)
    proofs = {json.dumps({k: v.get("proof") for k, v in r.items()}) for r in rows.values() if "error" not in r}
    print("all variants return the same proofs:", len(proofs) == 1)
    with open(os.path.join(ROOT, "gpurun_out", "ab_ntt.json"), "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
