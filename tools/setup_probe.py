#!/usr/bin/env python3
"""Where the seconds of a 2^k-constraint Groth16 trusted setup on the device go: circuit marshalling + QAP residency
(ps_qap_load_r1cs: CSR upload, transposes, Z-tree), then ps_g16_setup (exponents, fixed-base points, window tables).
    python tools/setup_probe.py [log_n]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import playsnark_b200 as ps
    from playsnark_b200 import synth
    from oracle import ps_oracle as O
    k = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    n = 1 << k
    be = ps.Backend(0)
    sq, wit = synth.sparse_circuit(n, 7, n // 2)
    smp = O.Sampler(99)
    toxic = tuple(smp.fr() for _ in range(5))
    res = {"log_n": k}
    for rep in range(2):
        t0 = time.perf_counter()
        sq._resident(be)
        be.sync()
        t1 = time.perf_counter()
        tr = ps.NewGroth16TrustedSetup(sq, backend=be, toxic=toxic, export=False)
        be.sync()
        t2 = time.perf_counter()
        res["run%d" % rep] = {"qap_resident_s": round(t1 - t0, 3), "g16_setup_s": round(t2 - t1, 3), "total_s": round(t2 - t0, 3)}
        tr.close(); sq.close()
    print("SETUP_PROBE " + json.dumps(res), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "setup_probe.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
