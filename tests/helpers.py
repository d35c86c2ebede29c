"""Shared helpers for the parity tests: oracle objects -> host-mirror (playsnark_b200.api) objects,
and synthetic circuits of the shapes BASELINE.json names."""
from __future__ import annotations

from oracle import ps_oracle as O
from playsnark_b200 import api


def mirror_qap(oq: O.QAP) -> api.QAP:
    return api.QAP(oq.nb_vars, oq.nb_io, oq.nb_gates, oq.left, oq.right, oq.out, oq.z)


def mirror_g16_setup(tr) -> api.Groth16Setup:
    g1, g2 = O.g1_compress, O.g2_compress
    return api.Groth16Setup(
        Alpha=g1(tr.Alpha), Beta=g1(tr.Beta), Delta=g1(tr.Delta), Xi=[g1(p) for p in tr.Xi],
        NioLP=[g1(p) for p in tr.NioLP], XiT=[g1(p) for p in tr.XiT], Beta2=g2(tr.Beta2), Delta2=g2(tr.Delta2),
        Xi2=[g2(p) for p in tr.Xi2], IoLP=[g1(p) for p in tr.IoLP], Gamma=g2(tr.Gamma))


def mirror_phgr13_ek(ek) -> api.PHGR13EvalKey:
    g1, g2 = O.g1_compress, O.g2_compress
    f = lambda k: [g1(p) for p in ek[k]]
    return api.PHGR13EvalKey(vs=f("vs"), ws=[g2(p) for p in ek["ws"]], ys=f("ys"), vas=f("vas"), was=f("was"),
                             yas=f("yas"), gsi=f("gsi"), vbs=f("vbs"), wbs=f("wbs"), ybs=f("ybs"))


def decode_phgr13(pp: api.PHGR13Proof) -> dict:
    d = {}
    for f in O.PHGR13_FIELDS:
        raw = getattr(pp, f)
        d[f] = O.g2_decompress(raw) if f == "wss" else O.g1_decompress(raw)
    return d


def squaring_chain(n: int, x0: int):
    """Config C2: x_{k+1} = x_k * x_k, n Mul gates, 1 input, 1 output (r1cs.go:148-152 shapes).
    Returns (R1CS, witness as Fr values)."""
    r = O.R1CS()
    r.new_input("x0")
    r.new_output("x%d" % n)
    for i in range(1, n):
        r.new_var("x%d" % i)
    for i in range(n):
        r.mul("x%d" % i, "x%d" % i, "x%d" % (i + 1))
    vals = {"const": 1}
    v = x0 % O.R
    for i in range(n + 1):
        vals["x%d" % i] = v
        v = v * v % O.R
    return r, [vals[nm] for nm in r.vars]


def mixed_circuit(n: int, seed: int, n_inputs: int):
    """n gates mixing Mul / Add / AddConst over earlier variables, many declared inputs so that the
    'last nbIO variables' segment (NioLP / PHGR13 mid) is large (SURVEY 8 hard parts).
    Returns (R1CS, witness as Fr values)."""
    import random
    rng = random.Random(seed)
    r = O.R1CS()
    names = []
    for i in range(n_inputs):
        r.new_input("in%d" % i); names.append("in%d" % i)
    r.new_output("out")
    for i in range(n - 1):
        r.new_var("t%d" % i)
    vals = {"const": 1}
    for nm in names:
        vals[nm] = rng.randrange(O.R)
    avail = list(names)
    for g in range(n):
        dst = "out" if g == n - 1 else "t%d" % g
        kind = rng.randrange(3)
        a, b = rng.choice(avail), rng.choice(avail)
        if kind == 0:
            r.mul(a, b, dst); vals[dst] = vals[a] * vals[b] % O.R
        elif kind == 1 and a != b:
            r.add(a, b, dst); vals[dst] = (vals[a] + vals[b]) % O.R
        else:
            k = rng.randrange(1, 100)
            r.add_const(a, k, dst); vals[dst] = (vals[a] + k) % O.R
        avail.append(dst)
    return r, [vals[nm] for nm in r.vars]
