#!/usr/bin/env python3
"""Generates tests/golden/*.json from the oracle (oracle/ps_oracle.py) with fixed seeds.

The reference itself cannot run here (no Go toolchain) and its tests pin no
curve-level bytes, so these vectors are produced by the oracle after it has passed its own anchors
(tests/test_oracle.py: public constants, the reference's integer KATs, pairing bilinearity, verifier
acceptance).  They freeze the oracle's outputs so that neither it nor the CUDA path can drift.
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ps_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
hx = lambda v: "%064x" % v


def readme_circuit():
    c = O.create_r1cs()
    w = O.create_witness(c)
    q = O.to_qap(c, fast=False)
    a, b, cc = q.compute_aggregate_poly(w)
    h = q.quotient(w)
    smp = O.Sampler(0)
    tr = O.groth16_setup(q, smp)
    r, s = smp.fr(), smp.fr()
    pr = O.groth16_prove(tr, q, w, r, s, faithful=True)
    g1, g2 = O.g1_compress, O.g2_compress
    st = O.phgr13_setup(q, O.Sampler(1))
    pp = O.phgr13_prove(st["EK"], q, w)
    return {
        "witness": w, "nb_vars": q.nb_vars, "nb_io": q.nb_io, "nb_gates": q.nb_gates,
        "left": [[hx(v) for v in p] for p in q.left], "right": [[hx(v) for v in p] for p in q.right],
        "out": [[hx(v) for v in p] for p in q.out], "z": [hx(v) for v in q.z],
        "a": [hx(v) for v in a], "b": [hx(v) for v in b], "c": [hx(v) for v in cc], "h": [hx(v) for v in h],
        "groth16": {
            "seed": 0, "r": hx(r), "s": hx(s),
            "Alpha": g1(tr.Alpha).hex(), "Beta": g1(tr.Beta).hex(), "Delta": g1(tr.Delta).hex(),
            "Beta2": g2(tr.Beta2).hex(), "Delta2": g2(tr.Delta2).hex(), "Gamma": g2(tr.Gamma).hex(),
            "Xi": [g1(p).hex() for p in tr.Xi], "Xi2": [g2(p).hex() for p in tr.Xi2],
            "XiT": [g1(p).hex() for p in tr.XiT], "NioLP": [g1(p).hex() for p in tr.NioLP],
            "IoLP": [g1(p).hex() for p in tr.IoLP],
            "A": g1(pr["A"]).hex(), "B": g2(pr["B"]).hex(), "C": g1(pr["C"]).hex(),
        },
        "phgr13": {
            "seed": 1,
            "ek": {k: [(g2 if k == "ws" else g1)(p).hex() for p in v] for k, v in st["EK"].items()},
            "proof": {f: (g2 if f == "wss" else g1)(pp[f]).hex() for f in O.PHGR13_FIELDS},
        },
    }


def msm_vectors():
    rng = random.Random(20261018)
    out = {}
    for grp, F, gen, comp in (("g1", O.F1, O.G1_GEN, O.g1_compress), ("g2", O.F2, O.G2_GEN, O.g2_compress)):
        n = 12
        ks = [rng.randrange(1, O.R) for _ in range(n)]
        pts = [O.pt_mul(F, k, gen) for k in ks]
        pts[5] = None            # point at infinity in the base set
        pts[7] = pts[6]          # repeated point
        sc = [rng.randrange(O.R) for _ in range(n)]
        sc[0], sc[1], sc[2], sc[3] = 0, 1, O.R - 1, 35
        res = O.msm_naive(F, sc, pts)
        out[grp] = {"points": [comp(p).hex() for p in pts], "scalars": [hx(v) for v in sc], "result": comp(res).hex()}
    return out


def ntt_vectors():
    rng = random.Random(7)
    v = [rng.randrange(O.R) for _ in range(16)]
    w = O.fr_root_of_unity(4)
    g = 5
    fwd = [sum(v[j] * pow(w, i * j, O.R) for j in range(16)) % O.R for i in range(16)]
    cos = [sum(v[j] * pow(g * pow(w, i, O.R), j, O.R) for j in range(16)) % O.R for i in range(16)]
    return {"input": [hx(x) for x in v], "omega": hx(w), "forward": [hx(x) for x in fwd], "coset": hx(g),
            "coset_forward": [hx(x) for x in cos]}


def constants():
    return {
        "p": "%x" % O.P, "r": "%x" % O.R,
        "g1_generator_compressed": O.g1_compress(O.G1_GEN).hex(),
        "g2_generator_compressed": O.g2_compress(O.G2_GEN).hex(),
        "g1_infinity_compressed": O.g1_compress(None).hex(),
        "g2_infinity_compressed": O.g2_compress(None).hex(),
        "g1_two_g": O.g1_compress(O.g1_mul(2)).hex(),
        "g2_two_g": O.g2_compress(O.g2_mul(2)).hex(),
        "root_of_unity_2_32": hx(O.fr_root_of_unity(32)),
    }


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    for name, fn in (("readme_circuit", readme_circuit), ("msm", msm_vectors), ("ntt", ntt_vectors), ("constants", constants)):
        with open(os.path.join(OUT, name + ".json"), "w") as f:
            json.dump(fn(), f, indent=1)
        print("wrote", name)
