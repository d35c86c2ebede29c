#!/usr/bin/env python3
"""MSM sweep for profiles/: G1 and G2, 2^12..2^max, random and witness-like scalar distributions.  Every `rand` row is
checked against (sum k_i s_i mod r)*G from the oracle (checker, untimed) before it is timed.
Usage: sweep_msm.py <g1_max_log> <g2_max_log> [tag]      -> gpurun_out/msm_sweep[_tag].json / .md"""
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
tag = ("_" + sys.argv[3]) if len(sys.argv) > 3 else ""
be = ps.Backend(0)
rng = np.random.default_rng(5)


def parity(group, ks, sc, got):
    """device result == (sum k_i s_i mod r) * G  (oracle's C dot product over Fr + one scalar multiplication)"""
    from oracle import c_oracle as CO, ps_oracle as O
    n = len(ks) // 32
    e = CO.fr_dot(np.frombuffer(ks, dtype=np.uint8).reshape(n, 32), np.frombuffer(sc, dtype=np.uint8).reshape(n, 32))
    want = O.g1_compress(O.g1_mul(e)) if group == L.PS_G1 else O.g2_compress(O.g2_mul(e))
    return bytes(got) == want


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
        ks = rand_scalars(n)
        bases = be.bases_from_scalars(group, ks, 0, -1)
        info = (__import__("ctypes").c_int * 4)()
        be.lib.ps_bases_info(bases.handle, info)
        kinds = ["rand"] + (["ones", "minus_one", "bits", "small64"] if log_n in (16, 20) else [])
        for kind in kinds:
            sc = kind_scalars(kind, n)
            got = be.msm(bases, sc)
            ok = parity(group, ks, sc, got) if kind == "rand" else None
            best = None
            for _ in range(3):
                be.msm(bases, sc)
                t = be.msm_timing()
                if best is None or t["total_ms"] < best["total_ms"]:
                    best = t
            row = dict(group=name, log_n=log_n, scalars=kind, c=info[0], W=info[1], T=info[2], **{k: round(v, 3) for k, v in best.items()})
            row["points_per_s"] = n / (max(best["total_ms"], 1e-6) * 1e-3)
            if ok is not None:
                row["parity"] = ok
            rows.append(row)
            print(row, flush=True)
        bases.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "msm_sweep%s.json" % tag), "w"), indent=1)
with open(os.path.join(ROOT, "gpurun_out", "msm_sweep%s.md" % tag), "w") as f:
    f.write("| group | points | scalars | c | W | sort ms | accumulate ms | merge ms | reduce ms | total ms | points/s | = oracle |\n")
    f.write("|---|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---|\n")
    for r in rows:
        f.write("| %s | 2^%d | %s | %d | %d | %.3f | %.3f | %.3f | %.3f | %.3f | %.3e | %s |\n" % (
            r["group"], r["log_n"], r["scalars"], r["c"], r["W"], r["sort_ms"], r["accumulate_ms"], r["combine_ms"],
            r["reduce_ms"], r["total_ms"], r["points_per_s"], {True: "yes", False: "NO", None: ""}[r.get("parity")]))
assert all(r.get("parity", True) for r in rows), "a sweep row differs from the oracle"
