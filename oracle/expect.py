"""CHECKER (test infrastructure, like everything under oracle/): exponent-level expectations for proofs of large
sparse circuits, in the style of the reference's own TestGroth16ProofGen (groth16_test.go:32-107) and
pinocchio_test.go:23-196 -- the proof elements recomputed in the exponent from the toxic waste, one scalar
multiplication per element, in O(nnz + n) field operations via the Lagrange basis at the toxic point.
Circuits are given as CSR triples (row_ptr, col, val) with attributes nbVars / nbIO / nbGates (api.SparseQAP shape).
"""
from __future__ import annotations

from . import ps_oracle as O

R = O.R


def lagrange_at(n: int, x: int):
    """l_j(x) for the domain {1..n}, j = 1..n (list index j-1), and z(x); O(n)."""
    z = 1
    for j in range(1, n + 1):
        z = z * (x - j) % R
    fact = [1] * (n + 1)
    for i in range(1, n + 1):
        fact[i] = fact[i - 1] * i % R
    den = []
    for j in range(1, n + 1):
        zp = fact[j - 1] * fact[n - j] % R
        if (n - j) & 1:
            zp = R - zp
        den.append((x - j) * zp % R)
    pref = [1] * (n + 1)
    for i, d in enumerate(den):
        pref[i + 1] = pref[i] * d % R
    inv = pow(pref[n], -1, R)
    out = [0] * n
    for i in range(n - 1, -1, -1):
        out[i] = z * (inv * pref[i] % R) % R
        inv = inv * den[i] % R
    return out, z


def eval_all(sq, lag):
    """u_i(x), v_i(x), w_i(x) for every variable i from the CSR rows and l_j(x)"""
    res = []
    for rp, col, val in (sq.left, sq.right, sq.out):
        acc = [0] * sq.nbVars
        for j in range(sq.nbGates):
            lj = lag[j]
            for k in range(rp[j], rp[j + 1]):
                acc[col[k]] = (acc[col[k]] + val[k] * lj) % R
        res.append(acc)
    return res


def aggregate_at(sq, wit, lag):
    """a(x), b(x), c(x): the witness-weighted sums at the toxic point"""
    ev = []
    for rp, col, val in (sq.left, sq.right, sq.out):
        e = 0
        for j in range(sq.nbGates):
            rowv = 0
            for k in range(rp[j], rp[j + 1]):
                rowv += val[k] * wit[col[k]]
            e = (e + rowv % R * lag[j]) % R
        ev.append(e)
    return ev


def groth16_expected(sq, wit, toxic, r: int, s: int):
    """toxic = (alpha, beta, delta, x, gamma).  Returns (A, B, C compressed bytes, (a(x), b(x), c(x)), z(x))."""
    alpha, beta, delta, x, _gamma = toxic
    n, m = sq.nbGates, sq.nbVars
    lag, zx = lagrange_at(n, x)
    u, v, w = eval_all(sq, lag)
    ax, bx, cx = aggregate_at(sq, wit, lag)
    dinv = pow(delta, -1, R)
    diff = m - sq.nbIO
    ea = (alpha + ax + r * delta) % R
    eb = (beta + bx + s * delta) % R
    ec = sum(wit[i] * ((beta * u[i] + alpha * v[i] + w[i]) * dinv % R) for i in range(diff, m)) % R
    ec = (ec + (ax * bx - cx) * dinv + s * ea + r * eb - r * s % R * delta) % R
    return O.g1_compress(O.g1_mul(ea)), O.g2_compress(O.g2_mul(eb)), O.g1_compress(O.g1_mul(ec)), (ax, bx, cx), zx


def phgr13_expected(sq, wit, toxic):
    """toxic = (s, av, aw, ay, rv, rw, beta, gamma).  Returns the eight proof elements as compressed bytes."""
    s, av, aw, ay, rv, rw, beta = toxic[:7]
    ry = rv * rw % R
    n, m = sq.nbGates, sq.nbVars
    diff = m - sq.nbIO
    lag, zs = lagrange_at(n, s)
    u, v, w = eval_all(sq, lag)
    ax, bx, cx = aggregate_at(sq, wit, lag)
    hs = (ax * bx - cx) * pow(zs, -1, R) % R
    dot = lambda vec: sum(wit[i] * vec[i] for i in range(diff, m)) % R
    vm, wm, ym = dot(u), dot(v), dot(w)
    g1 = lambda e: O.g1_compress(O.g1_mul(e % R))
    return {"hs": g1(hs), "vss": g1(rv * vm), "wss": O.g2_compress(O.g2_mul(rw * wm % R)), "yss": g1(ry * ym),
            "vass": g1(rv * vm * av), "wass": g1(rw * wm * aw), "yass": g1(ry * ym * ay),
            "gz": g1(beta * (rv * vm + rw * wm + ry * ym))}


def squaring_chain_r1cs(n: int, x0: int):
    """config C2's circuit as the oracle's dense R1CS (createR1CS-style builder calls, r1cs.go:148-152) + witness"""
    r = O.R1CS()
    r.new_input("x0")
    r.new_output("x%d" % n)
    for i in range(1, n):
        r.new_var("x%d" % i)
    for i in range(n):
        r.mul("x%d" % i, "x%d" % i, "x%d" % (i + 1))
    vals = {"const": 1}
    v = x0 % R
    for i in range(n + 1):
        vals["x%d" % i] = v
        v = v * v % R
    return r, [vals[nm] for nm in r.vars]
