"""Synthetic circuits of the shapes BASELINE.json names, built with the host mirror's own types.

  sparse_circuit   n gates of the reference's three shapes (Mul / Add / AddConst, r1cs.go:148-174) over earlier
                   variables, many declared inputs so that the "last nbIO variables" segment the provers sum over
                   (NioLP, groth16.go:173-179; PHGR13's mid range, pinochio.go:218-242) is large -- configs C3 / C5
  squaring_chain   x_{k+1} = x_k * x_k, n Mul gates, one input, one output -- config C2 (x0 = -1 is the only chain
                   that fits the reference's Value int)
No arithmetic beyond the witness values (Python ints mod r) happens here.
"""
from __future__ import annotations

import random

from . import api

R = api.R


def sparse_circuit(n: int, seed: int, n_inputs: int):
    """Returns (SparseQAP, witness as Fr ints); variable order [const, inputs..., out, intermediates...]
    (mergeVars, r1cs.go:132-144)."""
    rng = random.Random(seed)
    m = 1 + n_inputs + 1 + (n - 1)
    idx_out = 1 + n_inputs
    first_mid = idx_out + 1
    wit = [0] * m
    wit[0] = 1
    for i in range(1, 1 + n_inputs):
        wit[i] = rng.randrange(R)
    avail = list(range(1, 1 + n_inputs))
    L = ([0], [], []); Rm = ([0], [], []); Om = ([0], [], [])

    def push(mat, entries):
        for c, v in sorted(entries):
            mat[1].append(c); mat[2].append(v % R)
        mat[0].append(len(mat[1]))

    for g in range(n):
        dst = idx_out if g == n - 1 else first_mid + g
        kind = rng.randrange(3)
        a, b = rng.choice(avail), rng.choice(avail)
        if kind == 0:
            push(L, [(a, 1)]); push(Rm, [(b, 1)]); wit[dst] = wit[a] * wit[b] % R
        elif kind == 1 and a != b:
            push(L, [(a, 1), (b, 1)]); push(Rm, [(0, 1)]); wit[dst] = (wit[a] + wit[b]) % R
        else:
            k = rng.randrange(1, 100)
            push(L, [(0, k), (a, 1)]); push(Rm, [(0, 1)]); wit[dst] = (wit[a] + k) % R
        push(Om, [(dst, 1)])
        avail.append(dst)
    nb_io = 1 + n_inputs + 1
    return api.SparseQAP(m, nb_io, n, L, Rm, Om), wit


def squaring_chain(n: int, x0: int = -1):
    """Returns (api.R1CS, witness as Fr ints) for n repeated squarings of x0."""
    c = api.R1CS()
    c.NewInput("x0")
    c.NewOutput("x%d" % n)
    for i in range(1, n):
        c.NewVar("x%d" % i)
    for i in range(n):
        c.Mul("x%d" % i, "x%d" % i, "x%d" % (i + 1))
    vals = {"const": 1}
    v = x0 % R
    for i in range(n + 1):
        vals["x%d" % i] = v
        v = v * v % R
    return c, [vals[nm] for nm in c.vars]
