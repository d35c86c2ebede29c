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


# ---- large synthetic circuits kept sparse (configs C3 / C5 of BASELINE.json) ---------------------------
def sparse_circuit(n: int, seed: int, n_inputs: int):
    """n gates of the reference's three shapes (Mul / Add / AddConst, r1cs.go:148-174) over earlier
    variables, built directly in CSR.  Variable order [const, inputs..., out, intermediates...]
    (r1cs.go:132-144).  Returns (SparseQAP, witness as Fr ints)."""
    import random
    rng = random.Random(seed)
    m = 1 + n_inputs + 1 + (n - 1)
    idx_out = 1 + n_inputs
    first_mid = idx_out + 1
    wit = [0] * m
    wit[0] = 1
    for i in range(1, 1 + n_inputs):
        wit[i] = rng.randrange(O.R)
    avail = list(range(1, 1 + n_inputs))
    L = ([0], [], []); Rm = ([0], [], []); Om = ([0], [], [])

    def push(mat, entries):
        for c, v in sorted(entries):
            mat[1].append(c); mat[2].append(v % O.R)
        mat[0].append(len(mat[1]))

    for g in range(n):
        dst = idx_out if g == n - 1 else first_mid + g
        kind = rng.randrange(3)
        a, b = rng.choice(avail), rng.choice(avail)
        if kind == 0:
            push(L, [(a, 1)]); push(Rm, [(b, 1)]); wit[dst] = wit[a] * wit[b] % O.R
        elif kind == 1 and a != b:
            push(L, [(a, 1), (b, 1)]); push(Rm, [(0, 1)]); wit[dst] = (wit[a] + wit[b]) % O.R
        else:
            k = rng.randrange(1, 100)
            push(L, [(0, k), (a, 1)]); push(Rm, [(0, 1)]); wit[dst] = (wit[a] + k) % O.R
        push(Om, [(dst, 1)])
        avail.append(dst)
    nb_io = 1 + n_inputs + 1
    return api.SparseQAP(m, nb_io, n, L, Rm, Om), wit


def lagrange_at(n: int, x: int):
    """l_j(x) for the domain {1..n}, j = 1..n (list index j-1), and z(x); O(n)."""
    z = 1
    for j in range(1, n + 1):
        z = z * (x - j) % O.R
    fact = [1] * (n + 1)
    for i in range(1, n + 1):
        fact[i] = fact[i - 1] * i % O.R
    den = []
    for j in range(1, n + 1):
        zp = fact[j - 1] * fact[n - j] % O.R
        if (n - j) & 1:
            zp = O.R - zp
        den.append((x - j) * zp % O.R)
    # batch inversion
    pref = [1] * (n + 1)
    for i, d in enumerate(den):
        pref[i + 1] = pref[i] * d % O.R
    inv = pow(pref[n], -1, O.R)
    out = [0] * n
    for i in range(n - 1, -1, -1):
        out[i] = z * (inv * pref[i] % O.R) % O.R
        inv = inv * den[i] % O.R
    return out, z


def sparse_eval_all(sq, lag):
    """u_i(x), v_i(x), w_i(x) for every variable i (the per-variable QAP polynomials evaluated at
    the toxic point), from the CSR rows and l_j(x)."""
    res = []
    for rp, col, val in (sq.left, sq.right, sq.out):
        acc = [0] * sq.nbVars
        for j in range(sq.nbGates):
            lj = lag[j]
            for k in range(rp[j], rp[j + 1]):
                acc[col[k]] = (acc[col[k]] + val[k] * lj) % O.R
        res.append(acc)
    return res


def sparse_groth16_setup(be, sq, seed: int):
    """NewGroth16TrustedSetup (groth16.go:64-101) for a sparse circuit: exponents in Python, points by
    the device's fixed-base kernel.  Returns (api.Groth16Setup with uncompressed blobs, toxic dict)."""
    from playsnark_b200 import _lib as L
    smp = O.Sampler(seed)
    tw = {k: smp.fr() for k in ("Alpha", "Beta", "Delta", "X", "Gamma")}
    n, m = sq.nbGates, sq.nbVars
    x = tw["X"]
    lag, zx = lagrange_at(n, x)
    u, v, w = sparse_eval_all(sq, lag)
    dinv = pow(tw["Delta"], -1, O.R)
    diff = m - sq.nbIO
    nio = [(tw["Beta"] * u[i] + tw["Alpha"] * v[i] + w[i]) * dinv % O.R for i in range(diff, m)]
    pw = [1] * n
    for i in range(1, n):
        pw[i] = pw[i - 1] * x % O.R
    txd = zx * dinv % O.R
    A = L.PS_FMT_AFFINE
    pts = lambda grp, exps: be.bases_from_scalars(grp, exps).export(fmt=A, blob=True)
    one = lambda grp, e: be.bases_from_scalars(grp, [e]).export(fmt=A, blob=True)
    tr = api.Groth16Setup(
        Alpha=one(L.PS_G1, tw["Alpha"]), Beta=one(L.PS_G1, tw["Beta"]), Delta=one(L.PS_G1, tw["Delta"]),
        Xi=pts(L.PS_G1, pw), NioLP=pts(L.PS_G1, nio), XiT=pts(L.PS_G1, [p * txd % O.R for p in pw[:n - 1]]),
        Beta2=one(L.PS_G2, tw["Beta"]), Delta2=one(L.PS_G2, tw["Delta"]), Xi2=pts(L.PS_G2, pw), fmt=A)
    tw.update(lag=lag, zx=zx, nio=nio, diff=diff)
    return tr, tw


def sparse_groth16_expected(sq, wit, tw, r: int, s: int):
    """TestGroth16ProofGen (groth16_test.go:32-107) at scale: A, B, C recomputed in the exponent from
    the toxic waste in O(nnz + n), one scalar multiplication per element."""
    lag, n = tw["lag"], sq.nbGates
    ev = []
    for rp, col, val in (sq.left, sq.right, sq.out):
        e = 0
        for j in range(n):
            rowv = 0
            for k in range(rp[j], rp[j + 1]):
                rowv += val[k] * wit[col[k]]
            e = (e + rowv % O.R * lag[j]) % O.R
        ev.append(e)
    ax, bx, cx = ev
    dinv = pow(tw["Delta"], -1, O.R)
    ea = (tw["Alpha"] + ax + r * tw["Delta"]) % O.R
    eb = (tw["Beta"] + bx + s * tw["Delta"]) % O.R
    ec = sum(wv * e for wv, e in zip(wit[tw["diff"]:], tw["nio"])) % O.R
    ec = (ec + (ax * bx - cx) * dinv + s * ea + r * eb - r * s % O.R * tw["Delta"]) % O.R
    return O.g1_compress(O.g1_mul(ea)), O.g2_compress(O.g2_mul(eb)), O.g1_compress(O.g1_mul(ec)), (ax, bx, cx)


def sparse_phgr13_setup(be, sq, seed: int):
    """NewPHGR13TrustedSetup's evaluation key (pinochio.go:93-141) for a sparse circuit: exponents in
    Python, points by the device's fixed-base kernel.  Returns (api.PHGR13EvalKey, toxic dict)."""
    from playsnark_b200 import _lib as L
    smp = O.Sampler(seed)
    s = smp.fr()
    av, aw, ay = smp.fr(), smp.fr(), smp.fr()
    rv, rw = smp.fr(), smp.fr()
    ry = rv * rw % O.R
    beta = smp.fr()
    n, m = sq.nbGates, sq.nbVars
    lag, zs = lagrange_at(n, s)
    u, v, w = sparse_eval_all(sq, lag)
    diff = m - sq.nbIO
    pw = [1] * (n - 1)
    for i in range(1, n - 1):
        pw[i] = pw[i - 1] * s % O.R
    g1 = lambda exps: be.bases_from_scalars(L.PS_G1, exps).export()
    g2 = lambda exps: be.bases_from_scalars(L.PS_G2, exps).export()
    mid = range(diff, m)
    ek = api.PHGR13EvalKey(
        vs=g1([rv * u[i] % O.R for i in mid]), ws=g2([rw * v[i] % O.R for i in mid]), ys=g1([ry * w[i] % O.R for i in mid]),
        vas=g1([rv * u[i] * av % O.R for i in mid]), was=g1([rw * v[i] * aw % O.R for i in mid]),
        yas=g1([ry * w[i] * ay % O.R for i in mid]), gsi=g1(pw),
        vbs=g1([rv * u[i] * beta % O.R for i in mid]), wbs=g1([rw * v[i] * beta % O.R for i in mid]),
        ybs=g1([ry * w[i] * beta % O.R for i in mid]))
    tw = dict(s=s, av=av, aw=aw, ay=ay, rv=rv, rw=rw, ry=ry, beta=beta, u=u, v=v, w=w, diff=diff, lag=lag, zs=zs)
    return ek, tw


def sparse_phgr13_expected(sq, wit, tw):
    """the eight proof elements recomputed in the exponent (pinocchio_test.go:31-110 at scale)"""
    n, diff = sq.nbGates, tw["diff"]
    lag = tw["lag"]
    ev = []
    for rp, col, val in (sq.left, sq.right, sq.out):
        e = 0
        for j in range(n):
            rowv = 0
            for k in range(rp[j], rp[j + 1]):
                rowv += val[k] * wit[col[k]]
            e = (e + rowv % O.R * lag[j]) % O.R
        ev.append(e)
    ax, bx, cx = ev
    hs = (ax * bx - cx) * pow(tw["zs"], -1, O.R) % O.R
    dot = lambda vec: sum(wit[i] * vec[i] for i in range(diff, sq.nbVars)) % O.R
    vm, wm, ym = dot(tw["u"]), dot(tw["v"]), dot(tw["w"])
    rv, rw, ry, beta = tw["rv"], tw["rw"], tw["ry"], tw["beta"]
    g1 = lambda e: O.g1_compress(O.g1_mul(e % O.R))
    return {"hs": g1(hs), "vss": g1(rv * vm), "wss": O.g2_compress(O.g2_mul(rw * wm % O.R)), "yss": g1(ry * ym),
            "vass": g1(rv * vm * tw["av"]), "wass": g1(rw * wm * tw["aw"]), "yass": g1(ry * ym * tw["ay"]),
            "gz": g1(beta * (rv * vm + rw * wm + ry * ym))}
