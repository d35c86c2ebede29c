"""Parity cases shared by the host-emulation suite (CPU, small sizes) and the GPU suite (through the
product C ABI, larger sizes).  Each takes a playsnark_b200.Backend and compares with the oracle."""
from __future__ import annotations

import json
import os
import random

from oracle import ps_oracle as O
from playsnark_b200 import _lib as L, api
from tests import helpers as H

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gold(name):
    with open(os.path.join(GOLD, name + ".json")) as f:
        return json.load(f)


def _grp(group):
    if group == L.PS_G1:
        return O.F1, O.G1_GEN, O.g1_compress, O.g1_decompress
    return O.F2, O.G2_GEN, O.g2_compress, O.g2_decompress


def scalars_of(kind, n, rng):
    if kind == "rand":
        return [rng.randrange(O.R) for _ in range(n)]
    if kind == "ones":
        return [1] * n
    if kind == "neg":          # Value(-1).ToFieldElement() = r-1 (curve.go:17-19)
        return [O.R - 1] * n
    if kind == "small":        # witness-like: small signed ints
        return [rng.randrange(-40, 41) % O.R for _ in range(n)]
    if kind == "zero":
        return [0] * n
    if kind == "edge":
        base = [0, 1, 2, O.R - 1, O.R - 2, (O.R - 1) // 2, (O.R + 1) // 2, 1 << 254, (1 << 128) - 1]
        return [base[i % len(base)] for i in range(n)]
    raise ValueError(kind)


def msm_exponent_check(be, group, n, kind="rand", window_bits=0, tables=1, seed=0):
    """bases k_i*G built on the device; the MSM must equal (sum k_i s_i)*G (exponent-level check in
    the style of groth16_test.go:32-107 -- needs one scalar-mul, so it scales to any n)."""
    F, gen, comp, _ = _grp(group)
    rng = random.Random((seed << 8) ^ n ^ (group << 30))
    ks = [rng.randrange(1, O.R) for _ in range(n)]
    sc = scalars_of(kind, n, rng)
    bases = be.bases_from_scalars(group, ks, window_bits, tables)
    got = be.msm(bases, sc)
    want = comp(O.pt_mul(F, sum(k * s for k, s in zip(ks, sc)) % O.R, gen))
    assert got == want, (group, n, kind, window_bits, tables)
    bases.close()


def msm_exponent_check_big(be, group, log_n, window_bits=0, tables=-1, seed=0, resident=True):
    """the same check at the benchmark's sizes (2^24 G1 / 2^20 G2): inputs generated with numpy as
    bench.py does, the expectation from the oracle's C dot product over Fr; with `tables` = -1 the
    bases carry all window tables and the library picks the window (c = 22 and the two-pass scatter
    at 2^24).  Both the host-scalar call (ps_msm) and the device-resident one (ps_msm_device +
    ps_msm_combine) are checked."""
    import ctypes as C
    import numpy as np
    from oracle import c_oracle as CO
    F, gen, comp, _ = _grp(group)
    n = 1 << log_n
    rng = np.random.default_rng(seed * 1000 + log_n * 4 + group)

    def scalars():
        a = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
        a[:, 0] &= 0x3F
        return a
    ks, sc = scalars(), scalars()
    bases = be.bases_from_scalars(group, ks.tobytes(), window_bits, tables)
    want = comp(O.pt_mul(F, CO.fr_dot(ks, sc), gen))
    del ks
    out = C.create_string_buffer(48 if group == L.PS_G1 else 96)
    st = be.lib.ps_msm(be.ctx, bases.handle, C.c_void_p(sc.ctypes.data), n, out)
    be._check(st)
    assert out.raw == want, (group, log_n, "ps_msm")
    if resident:
        import torch
        limbs = np.ascontiguousarray(sc[:, ::-1]).view("<u4").reshape(-1, 8)
        d_sc = torch.from_numpy(limbs.view(np.int32)).cuda()
        d_part = torch.zeros(192 if group == L.PS_G1 else 384, dtype=torch.uint8, device="cuda")
        be._check(be.lib.ps_msm_device(be.ctx, bases.handle, 0, C.c_void_p(d_sc.data_ptr()), n, C.c_void_p(d_part.data_ptr())))
        be._check(be.lib.ps_msm_combine(be.ctx, group, C.c_void_p(d_part.data_ptr()), 1, out))
        assert out.raw == want, (group, log_n, "ps_msm_device")
    info = (C.c_int * 4)()
    be._check(be.lib.ps_bases_info(bases.handle, info))
    bases.close()
    return info[0], info[1], info[2]


def msm_vs_naive(be, group, n, seed=1):
    """against the reference's own algorithm (BlindEval: one scalar-mul per term)."""
    F, gen, comp, _ = _grp(group)
    rng = random.Random(seed)
    pts = [O.pt_mul(F, rng.randrange(1, O.R), gen) for _ in range(n)]
    sc = [rng.randrange(O.R) for _ in range(n)]
    bases = be.load_bases(group, [comp(p) for p in pts])
    assert be.msm(bases, sc) == comp(O.msm_naive(F, sc, pts))


def msm_golden(be):
    g = gold("msm")
    for name, group in (("g1", L.PS_G1), ("g2", L.PS_G2)):
        v = g[name]
        bases = be.load_bases(group, [bytes.fromhex(p) for p in v["points"]])
        assert be.msm(bases, [int(s, 16) for s in v["scalars"]]).hex() == v["result"]
        # the same bases through the uncompressed format
        aff = bases.export(fmt=L.PS_FMT_AFFINE)
        b2 = be.load_bases(group, aff, fmt=L.PS_FMT_AFFINE)
        assert b2.export() == [bytes.fromhex(p) for p in v["points"]]
        assert be.msm(b2, [int(s, 16) for s in v["scalars"]]).hex() == v["result"]


def msm_errors(be):
    import pytest
    bases = be.bases_from_scalars(L.PS_G1, [1, 2, 3])
    with pytest.raises(ValueError):      # BlindEval length panic, algebra.go:350-352
        be.msm(bases, [1, 2])
    with pytest.raises(api.L.PlaysnarkError) as e:   # non-canonical scalar
        be.msm(bases, b"\xff" * 96)
    assert e.value.status == L.PS_ERR_ENCODING
    with pytest.raises(api.L.PlaysnarkError) as e:   # x not on the curve
        be.load_bases(L.PS_G1, [bytes([0x80]) + bytes(46) + b"\x01"])
    assert e.value.status == L.PS_ERR_ENCODING
    # empty MSM = identity
    empty = be.load_bases(L.PS_G1, b"")
    assert be.msm(empty, []) == O.g1_compress(None)
    # points ON the curve but outside the prime-order subgroup are rejected at load, as kilic's
    # FromCompressed does for the reference's keys (the MSM's k -> r-k folding relies on r*P = O)
    for group, F, comp in ((L.PS_G1, O.F1, O.g1_compress), (L.PS_G2, O.F2, O.g2_compress)):
        x = 1
        while True:
            x += 1
            xx = x if group == L.PS_G1 else (x, 1)
            rhs = F.add(F.mul(F.sqr(xx), xx), F.b)
            y = O.fp_sqrt(rhs) if group == L.PS_G1 else O.fp2_sqrt(rhs)
            if y is not None and not O.in_subgroup(F, (xx, y)):
                break
        enc = comp((xx, y))
        with pytest.raises(api.L.PlaysnarkError) as e:
            be.load_bases(group, [enc])
        assert e.value.status == L.PS_ERR_ENCODING
        be.set_option("subgroup_check", 0)       # a caller that vouches for its key may skip the check
        try:
            assert be.load_bases(group, [enc]).export() == [enc]
        finally:
            be.set_option("subgroup_check", 1)
    # non-canonical encodings of infinity
    for bad in (bytes([0xC0]) + bytes(46) + b"\x01", bytes([0xE0]) + bytes(47)):
        with pytest.raises(api.L.PlaysnarkError) as e:
            be.load_bases(L.PS_G1, [bad])
        assert e.value.status == L.PS_ERR_ENCODING


def codec_roundtrip(be, n=24):
    rng = random.Random(3)
    for group in (L.PS_G1, L.PS_G2):
        F, gen, comp, _ = _grp(group)
        pts = [O.pt_mul(F, rng.randrange(1, O.R), gen) for _ in range(n)] + [None]
        enc = [comp(p) for p in pts]
        b = be.load_bases(group, enc)
        assert b.export() == enc
        affb = O.g1_affine_bytes if group == L.PS_G1 else O.g2_affine_bytes
        assert b.export(fmt=L.PS_FMT_AFFINE) == [affb(p) for p in pts]
    c = gold("constants")
    g = be.bases_from_scalars(L.PS_G1, [1, 2, 0])
    assert [x.hex() for x in g.export()] == [c["g1_generator_compressed"], c["g1_two_g"], c["g1_infinity_compressed"]]
    g = be.bases_from_scalars(L.PS_G2, [1, 2, 0])
    assert [x.hex() for x in g.export()] == [c["g2_generator_compressed"], c["g2_two_g"], c["g2_infinity_compressed"]]


def msm_linearity(be, n):
    """MSM(s+t) = MSM(s) + MSM(t), MSM(k*s) = k*MSM(s) on a fixed base set."""
    rng = random.Random(42)
    bases = be.bases_from_scalars(L.PS_G1, [rng.randrange(1, O.R) for _ in range(n)])
    s = [rng.randrange(O.R) for _ in range(n)]
    t = [rng.randrange(O.R) for _ in range(n)]
    k = rng.randrange(O.R)
    ms, mt = O.g1_decompress(be.msm(bases, s)), O.g1_decompress(be.msm(bases, t))
    mst = O.g1_decompress(be.msm(bases, [(a + b) % O.R for a, b in zip(s, t)]))
    assert mst == O.g1_add(ms, mt)
    assert O.g1_decompress(be.msm(bases, [a * k % O.R for a in s])) == O.g1_mul(k, ms)


def ntt_cases(be, max_log=7):
    g = gold("ntt")
    v = [int(x, 16) for x in g["input"]]
    assert be.ntt(v) == [int(x, 16) for x in g["forward"]]
    assert be.ntt(v, coset=int(g["coset"], 16)) == [int(x, 16) for x in g["coset_forward"]]
    rng = random.Random(11)
    for log_n in range(0, max_log + 1):
        n = 1 << log_n
        v = [rng.randrange(O.R) for _ in range(n)]
        got = be.ntt(v)
        if n <= 64:
            w = O.fr_root_of_unity(log_n)
            assert got == [sum(v[j] * pow(w, i * j, O.R) for j in range(n)) % O.R for i in range(n)]
        assert be.ntt(got, inverse=True) == v
        c = rng.randrange(2, O.R)
        assert be.ntt(be.ntt(v, coset=c), inverse=True, coset=c) == v


def ntt_properties(be, log_n):
    """size-independent properties: evaluation at a few points, linearity, round trip."""
    n = 1 << log_n
    rng = random.Random(log_n)
    v = [rng.randrange(O.R) for _ in range(n)]
    u = [rng.randrange(O.R) for _ in range(n)]
    fv, fu = be.ntt(v), be.ntt(u)
    w = O.fr_root_of_unity(log_n)
    for i in (0, 1, 2, n // 2, n - 1, rng.randrange(n)):
        assert fv[i] == O.poly_eval(v, pow(w, i, O.R))
    k = rng.randrange(O.R)
    assert be.ntt([(a + k * b) % O.R for a, b in zip(v, u)]) == [(a + k * b) % O.R for a, b in zip(fv, fu)]
    assert be.ntt(fv, inverse=True) == v


def readme_quotient(be):
    g = gold("readme_circuit")
    I = lambda xs: [int(x, 16) for x in xs]
    q = api.QAP(g["nb_vars"], g["nb_io"], g["nb_gates"], [I(p) for p in g["left"]], [I(p) for p in g["right"]],
                [I(p) for p in g["out"]], I(g["z"]))
    h, (a, b, c) = api.Quotient(q, g["witness"], backend=be, return_abc=True)
    assert h == I(g["h"]) and a == I(g["a"]) and b == I(g["b"]) and c == I(g["c"])
    assert len(h) == g["nb_gates"] - 1          # TestGroth16TrustedSetup, groth16_test.go:16-19
    import pytest
    bad = list(g["witness"]); bad[3] += 1
    with pytest.raises(ArithmeticError, match="apocalypse"):   # qap.go:158-160
        api.Quotient(q, bad, backend=be)
    with pytest.raises(ValueError):                              # sanityCheck, qap.go:177-189
        api.Quotient(q, g["witness"][:-1], backend=be)
    return q


def quotient_vs_div2(be, n, seed=0, circuit="chain"):
    rng = random.Random(seed)
    if circuit == "chain":
        r, w = H.squaring_chain(n, rng.randrange(O.R))
    else:
        r, w = H.mixed_circuit(n, seed, max(1, n // 2))
    oq = O.to_qap(r)
    q = H.mirror_qap(oq)
    h, (a, b, c) = api.Quotient(q, w, backend=be, return_abc=True)
    oa, ob, oc = oq.compute_aggregate_poly(w)
    assert (a, b, c) == (oa, ob, oc)
    assert h == oq.quotient(w)        # Poly.Div2 restatement
    import pytest
    bad = list(w); bad[-1] = (bad[-1] + 1) % O.R
    with pytest.raises(ArithmeticError):
        api.Quotient(q, bad, backend=be)
    return r, w, oq, q


def readme_groth16(be):
    g = gold("readme_circuit")
    I = lambda xs: [int(x, 16) for x in xs]
    q = api.QAP(g["nb_vars"], g["nb_io"], g["nb_gates"], [I(p) for p in g["left"]], [I(p) for p in g["right"]],
                [I(p) for p in g["out"]], I(g["z"]))
    k = g["groth16"]
    B = bytes.fromhex
    tr = api.Groth16Setup(Alpha=B(k["Alpha"]), Beta=B(k["Beta"]), Delta=B(k["Delta"]), Xi=[B(x) for x in k["Xi"]],
                          NioLP=[B(x) for x in k["NioLP"]], XiT=[B(x) for x in k["XiT"]], Beta2=B(k["Beta2"]),
                          Delta2=B(k["Delta2"]), Xi2=[B(x) for x in k["Xi2"]])
    pr = api.Groth16Prove(tr, q, g["witness"], int(k["r"], 16), int(k["s"], 16), backend=be, want_h=True)
    assert (pr.A.hex(), pr.B.hex(), pr.C.hex()) == (k["A"], k["B"], k["C"])
    assert pr.h == I(g["h"])
    assert pr.tp.R == int(k["r"], 16) and pr.tp.S == int(k["s"], 16)
    # the same with the witness marshalled into page-locked host memory (ps_host_alloc)
    hb = api.HostBuffer(be, g["witness"])
    pr1 = api.Groth16Prove(tr, q, hb, int(k["r"], 16), int(k["s"], 16), backend=be)
    assert (pr1.A.hex(), pr1.B.hex(), pr1.C.hex()) == (k["A"], k["B"], k["C"])
    assert api.Quotient(q, hb, backend=be) == I(g["h"])
    hb.close()
    # fresh randomness path (groth16.go:148,158): proof must still verify
    c = O.create_r1cs(); w = O.create_witness(c); oq = O.to_qap(c)
    otr = O.groth16_setup(oq, O.Sampler(0))
    pr2 = api.Groth16Prove(tr, q, w, backend=be)
    dec = {"A": O.g1_decompress(pr2.A), "B": O.g2_decompress(pr2.B), "C": O.g1_decompress(pr2.C)}
    assert O.groth16_verify(otr, oq, dec, w[:oq.nb_vars - oq.nb_io])          # TestGroth16Verify
    ea, eb, ec = O.groth16_expected_from_toxic(otr, oq, w, pr2.tp.R, pr2.tp.S)  # TestGroth16ProofGen
    assert (dec["A"], dec["B"], dec["C"]) == (ea, eb, ec)


def readme_phgr13(be):
    g = gold("readme_circuit")
    I = lambda xs: [int(x, 16) for x in xs]
    q = api.QAP(g["nb_vars"], g["nb_io"], g["nb_gates"], [I(p) for p in g["left"]], [I(p) for p in g["right"]],
                [I(p) for p in g["out"]], I(g["z"]))
    k = g["phgr13"]
    ek = api.PHGR13EvalKey(**{name: [bytes.fromhex(x) for x in v] for name, v in k["ek"].items()})
    pp = api.PHGR13Prove(ek, q, g["witness"], backend=be, want_h=True)
    for f in O.PHGR13_FIELDS:
        assert getattr(pp, f).hex() == k["proof"][f], f
    assert pp.h == I(g["h"])
    # verifier acceptance + the mutation rejects of pinocchio_test.go:243-276
    c = O.create_r1cs(); w = O.create_witness(c); oq = O.to_qap(c)
    st = O.phgr13_setup(oq, O.Sampler(1))
    dec = H.decode_phgr13(pp)
    io = w[:oq.nb_vars - oq.nb_io]
    assert O.phgr13_verify(st["VK"], oq, dec, io)
    for f in ("hs", "vss", "vass", "yass", "gz"):
        bad = dict(dec); bad[f] = O.g1_add(dec[f], O.G1_GEN)
        assert not O.phgr13_verify(st["VK"], oq, bad, io), f


def groth16_circuit(be, n, seed, circuit="mixed", verify=True):
    """full prove on a synthetic circuit: bit-exact vs the oracle prover (closed form), exponent-level
    recomputation from the toxic waste, and the oracle's pairing verifier."""
    rng = random.Random(seed)
    if circuit == "chain":
        r, w = H.squaring_chain(n, rng.choice([O.R - 1, rng.randrange(O.R)]))
    else:
        r, w = H.mixed_circuit(n, seed, max(1, n // 2))
    oq = O.to_qap(r)
    smp = O.Sampler(seed)
    tr = O.groth16_setup(oq, smp)
    rr, ss = smp.fr(), smp.fr()
    pr = api.Groth16Prove(H.mirror_g16_setup(tr), H.mirror_qap(oq), w, rr, ss, backend=be, want_h=True)
    want = O.groth16_prove(tr, oq, w, rr, ss)
    assert pr.h == want["h"]
    assert pr.A == O.g1_compress(want["A"]) and pr.B == O.g2_compress(want["B"]) and pr.C == O.g1_compress(want["C"])
    ea, eb, ec = O.groth16_expected_from_toxic(tr, oq, w, rr, ss)
    assert (want["A"], want["B"], want["C"]) == (ea, eb, ec)
    if verify:
        assert O.groth16_verify(tr, oq, want, w[:oq.nb_vars - oq.nb_io])


def phgr13_circuit(be, n, seed, verify=True):
    r, w = H.mixed_circuit(n, seed, max(1, n // 2))
    oq = O.to_qap(r)
    st = O.phgr13_setup(oq, O.Sampler(seed))
    pp = api.PHGR13Prove(H.mirror_phgr13_ek(st["EK"]), H.mirror_qap(oq), w, backend=be, want_h=True)
    want = O.phgr13_prove(st["EK"], oq, w)
    for f in O.PHGR13_FIELDS:
        enc = O.g2_compress if f == "wss" else O.g1_compress
        assert getattr(pp, f) == enc(want[f]), f
    assert pp.h == want["h"]
    if verify:
        assert O.phgr13_verify(st["VK"], oq, want, w[:oq.nb_vars - oq.nb_io])


def sparse_quotient_vs_dense(be, n, seed):
    """the sparse-R1CS path (SpMV + interpolation on {1..n}) against the oracle's Interpolate / Div2"""
    r, w = H.mixed_circuit(n, seed, max(1, n // 2))
    oq = O.to_qap(r)
    sq = api.SparseQAP.from_dense_rows(len(r.vars), r.nb_io(), r.left, r.right, r.out)
    h, (a, b, c) = api.Quotient(sq, w, backend=be, return_abc=True)
    assert (a, b, c) == tuple(oq.compute_aggregate_poly(w))
    assert h == oq.quotient(w)
    import pytest
    bad = list(w); bad[-1] = (bad[-1] + 1) % O.R
    with pytest.raises(ArithmeticError, match="apocalypse"):
        api.Quotient(sq, bad, backend=be)
    # dense and sparse forms of the same circuit give the same proof
    smp = O.Sampler(seed)
    tr = O.groth16_setup(oq, smp)
    rr, ss = smp.fr(), smp.fr()
    mtr = H.mirror_g16_setup(tr)
    p1 = api.Groth16Prove(mtr, H.mirror_qap(oq), w, rr, ss, backend=be)
    p2 = api.Groth16Prove(mtr, sq, w, rr, ss, backend=be)
    assert (p1.A, p1.B, p1.C) == (p2.A, p2.B, p2.C)
    want = O.groth16_prove(tr, oq, w, rr, ss)
    assert p2.A == O.g1_compress(want["A"]) and p2.B == O.g2_compress(want["B"]) and p2.C == O.g1_compress(want["C"])


def groth16_sparse_exponent_check(be, log_n, seed, n=None):
    """full Groth16 prove on a sparse synthetic circuit of 2^log_n gates (configs C3/C5): the proof
    must equal the exponent-level recomputation from the toxic waste, and h must satisfy
    h(x) z(x) = a(x) b(x) - c(x) at the toxic point."""
    n = n or (1 << log_n)
    sq, wit = H.sparse_circuit(n, seed, n // 2)
    tr, tw = H.sparse_groth16_setup(be, sq, seed)
    smp = O.Sampler(seed + 1000)
    r, s = smp.fr(), smp.fr()
    pr = api.Groth16Prove(tr, sq, wit, r, s, backend=be, want_h=True)
    A, B, Cc, (ax, bx, cx) = H.sparse_groth16_expected(sq, wit, tw, r, s)
    assert len(pr.h) == n - 1
    assert O.poly_eval(pr.h, tw["X"]) * tw["zx"] % O.R == (ax * bx - cx) % O.R
    assert pr.A == A and pr.B == B and pr.C == Cc
    import pytest
    bad = list(wit); bad[-1] = (bad[-1] + 1) % O.R
    with pytest.raises(ArithmeticError, match="apocalypse"):
        api.Groth16Prove(tr, sq, bad, r, s, backend=be)


def sharded_steps_recombine(be, log_n, parts, fake_world, seed, device="cpu", n=None):
    """The entry points of the multi-GPU Groth16 flow (dist.py), driven from ONE process: the subtrees
    of ps_qap_interp_part + ps_qap_interp_finish must reproduce ps_qap_aggregate_one's coefficients,
    ps_g16_scalars_ab / ps_g16_h_from_ab the scalar vectors of ps_g16_scalars, and the partial MSMs of
    `fake_world` uneven shards, added by ps_g16_combine, the proof of ps_g16_prove -- all bit for bit."""
    import ctypes as C
    import torch
    from playsnark_b200 import dist as D
    lib = be.lib
    n = n or (1 << log_n)
    n_leaves = D._tree_leaves(n)       # the subtree roots are sized by the tree's leaves (power of two >= n)
    sq, wit = H.sparse_circuit(n, seed, n // 2)
    tr, tw = H.sparse_groth16_setup(be, sq, seed)
    smp = O.Sampler(seed + 77)
    r, s = smp.fr(), smp.fr()
    pr = api.Groth16Prove(tr, sq, wit, r, s, backend=be)
    kh, qh = tr._resident(be), sq._resident(be)
    wb, rb, sb = api._fr_bytes(wit), api._fr_bytes([r]), api._fr_bytes([s])
    new = lambda rows: torch.zeros((rows, 8), dtype=torch.int32, device=device)
    ptr = lambda t: C.c_void_p(t.data_ptr())
    nio = sq.nbIO
    nA, nC, nB = (int(lib.ps_g16_scalar_count(kh, w)) for w in (0, 1, 2))
    head = nio + n - 1
    # reference scalar vectors and coefficients from the single-GPU entry points
    refA, refC, refB = new(nA), new(nC), new(nB)
    be._check(lib.ps_g16_scalars(be.ctx, kh, qh, wb, rb, sb, ptr(refA), ptr(refC), ptr(refB)))
    status = torch.zeros(1, dtype=torch.int32, device=device)
    bufA, bufC, bufB = new(nA), new(nC), new(nB)
    coef = []
    for which in (0, 1):
        want = new(n)
        be._check(lib.ps_qap_aggregate_one(be.ctx, qh, wb, which, ptr(want)))
        if parts == 1:
            got = new(n)
            be._check(lib.ps_qap_interp_part(be.ctx, qh, wb, which, 0, 1, ptr(got), ptr(bufC), ptr(status)))
        else:
            rows = 2 * n_leaves // parts
            e_all = new(parts * rows)
            for part in range(parts):
                be._check(lib.ps_qap_interp_part(be.ctx, qh, wb, which, part, parts, ptr(e_all[part * rows:(part + 1) * rows]),
                                                 ptr(bufC), ptr(status)))
            got = new(n)
            be._check(lib.ps_qap_interp_finish(be.ctx, qh, parts, ptr(e_all), ptr(got)))
        be.sync()
        assert torch.equal(got, want), ("coefficients", which)
        coef.append(got)
    # witness uploaded in two slices (ps_fr_upload) and used from device memory: same subtree root
    m = sq.nbVars
    wdev = new(m)
    half = m // 2
    be._check(lib.ps_fr_upload(be.ctx, wb[:32 * half], half, ptr(wdev), ptr(status)))
    be._check(lib.ps_fr_upload(be.ctx, wb[32 * half:], m - half, ptr(wdev[half:]), ptr(status)))
    rows = 2 * n_leaves // parts if parts > 1 else n
    via_host, via_dev, nio_dev = new(rows), new(rows), new(max(1, nio))
    be._check(lib.ps_qap_interp_part(be.ctx, qh, wb, 1, parts - 1, parts, ptr(via_host), None, ptr(status)))
    be._check(lib.ps_qap_interp_part_dev(be.ctx, qh, ptr(wdev), 1, parts - 1, parts, ptr(via_dev), ptr(nio_dev), ptr(status)))
    be.sync()
    assert torch.equal(via_host, via_dev) and torch.equal(nio_dev[:nio], bufC[:nio])
    be._check(lib.ps_g16_scalars_ab(be.ctx, kh, rb, sb, ptr(coef[0]), ptr(coef[1]), ptr(bufA), ptr(bufB), ptr(bufC[head:])))
    be._check(lib.ps_g16_h_from_ab(be.ctx, qh, ptr(coef[0]), ptr(coef[1]), ptr(bufC[nio:head])))
    be.sync()
    assert int(status[0]) == 0
    assert torch.equal(bufA, refA) and torch.equal(bufB, refB) and torch.equal(bufC, refC)
    # uneven shards of the four MSM pieces, one 976-byte record per fake rank
    weights = [0.4] + [1.0] * (fake_world - 1)
    recs = torch.zeros((fake_world, 976), dtype=torch.uint8, device=device)
    spans = lambda cnt: D.weighted_ranges(cnt, weights)
    rA, rB, rT, rH = spans(nA), spans(nB), spans(nC - head), spans(head)

    def partials(sp):
        out = torch.zeros(768, dtype=torch.uint8, device=device)
        first = (C.c_size_t * 3)(*[lo for lo, _ in sp])
        cnt = (C.c_size_t * 3)(*[hi - lo for lo, hi in sp])
        views = [buf[lo:hi] if hi > lo else buf for buf, (lo, hi) in zip((bufA, bufC, bufB), sp)]
        be._check(lib.ps_g16_msm_partials(be.ctx, kh, ptr(views[0]), ptr(views[1]), ptr(views[2]), first, cnt, ptr(out)))
        be.sync()
        return out

    for k in range(fake_world):
        early = partials([rA[k], (head + rT[k][0], head + rT[k][1]), rB[k]])
        late = partials([(0, 0), rH[k], (0, 0)])
        recs[k, :768] = early
        recs[k, 768:960] = late[192:384]
    assert sum(hi - lo for lo, hi in rA) == nA and rA[0][0] == 0 and rA[-1][1] == nA
    oA, oB, oC = C.create_string_buffer(48), C.create_string_buffer(96), C.create_string_buffer(48)
    be._check(lib.ps_g16_combine(be.ctx, ptr(recs), fake_world, 976, oA, oB, oC))
    assert (oA.raw, oB.raw, oC.raw) == (pr.A, pr.B, pr.C)
    # a broken witness sets the remainder bit of the device status word on the rank that owns the gate
    bad = list(wit); bad[-1] = (bad[-1] + 1) % O.R
    st2 = torch.zeros(1, dtype=torch.int32, device=device)
    hit = 0
    for part in range(parts):
        st2.zero_()
        rows = 2 * n_leaves // parts if parts > 1 else n
        sink = new(rows)                                       # kept alive across the call
        be._check(lib.ps_qap_interp_part(be.ctx, qh, api._fr_bytes(bad), 0, part, parts, ptr(sink), None, ptr(st2)))
        be.sync()
        hit += int(st2[0]) & 2
    assert hit >= 2
    st3 = torch.zeros(1, dtype=torch.int32, device=device)     # a non-canonical scalar sets the encoding bit
    sink = new(1)
    be._check(lib.ps_fr_upload(be.ctx, b"\xff" * 32, 1, ptr(sink), ptr(st3)))
    be.sync()
    assert int(st3[0]) & 1


def config_c2(be, n=1 << 10, timings=None):
    """BASELINE configs[1]: repeated-squaring R1CS with 2^10 multiplication gates (dense QAP of
    3*1026*1024 coefficients), Groth16 and PHGR13 prove; int witness x0 = -1 (the only chain that fits
    the reference's Value int) -- a full-width negative scalar r-1 repeated."""
    r, w = H.squaring_chain(n, O.R - 1)
    assert set(w) <= {1, O.R - 1}
    oq = O.to_qap(r)
    q = H.mirror_qap(oq)
    # Groth16: setup by scalar exponents on the device's fixed-base kernel, checked in the exponent
    smp = O.Sampler(2)
    tw = {k: smp.fr() for k in ("Alpha", "Beta", "Delta", "X", "Gamma")}
    x = tw["X"]
    pw = [pow(x, i, O.R) for i in range(n)]
    dinv = pow(tw["Delta"], -1, O.R)
    diff = oq.nb_vars - oq.nb_io
    nio = [O.linear_poly_for_var(oq, i, x, tw["Alpha"], tw["Beta"]) * dinv % O.R for i in range(diff, oq.nb_vars)]
    txd = O.poly_eval(oq.z, x) * dinv % O.R
    g1 = lambda exps: be.bases_from_scalars(L.PS_G1, exps).export()
    g2 = lambda exps: be.bases_from_scalars(L.PS_G2, exps).export()
    tr = api.Groth16Setup(Alpha=g1([tw["Alpha"]])[0], Beta=g1([tw["Beta"]])[0], Delta=g1([tw["Delta"]])[0], Xi=g1(pw),
                          NioLP=g1(nio), XiT=g1([p * txd % O.R for p in pw[:n - 1]]), Beta2=g2([tw["Beta"]])[0],
                          Delta2=g2([tw["Delta"]])[0], Xi2=g2(pw))
    rr, ss = smp.fr(), smp.fr()
    pr = api.Groth16Prove(tr, q, w, rr, ss, backend=be, want_h=True)
    h = oq.quotient(w)                                   # Poly.Div2 restatement, n = 2^10
    assert pr.h == h and len(h) == n - 1
    a, b, c = oq.compute_aggregate_poly(w)
    ax, bx = O.poly_eval(a, x), O.poly_eval(b, x)
    ea = (tw["Alpha"] + ax + rr * tw["Delta"]) % O.R
    eb = (tw["Beta"] + bx + ss * tw["Delta"]) % O.R
    ec = (sum(wv * e for wv, e in zip(w[diff:], nio)) + O.poly_eval(h, x) * txd + ss * ea + rr * eb
          - rr * ss % O.R * tw["Delta"]) % O.R
    assert pr.A == O.g1_compress(O.g1_mul(ea)) and pr.B == O.g2_compress(O.g2_mul(eb)) and pr.C == O.g1_compress(O.g1_mul(ec))
    # PHGR13 on the same QAP: evaluation key from the oracle (prover side only), proof vs the oracle prover
    st = O.phgr13_setup(oq, O.Sampler(3), with_vk=False)
    pp = api.PHGR13Prove(H.mirror_phgr13_ek(st["EK"]), q, w, backend=be, want_h=True)
    assert pp.h == h
    assert pp.hs == O.g1_compress(O.g1_mul(O.poly_eval(h, st["t"]["s"])))      # pinocchio_test.go:31-44
    fe = w[diff:]
    for f, key, F, comp in (("vss", "vs", O.F1, O.g1_compress), ("yss", "ys", O.F1, O.g1_compress),
                            ("vass", "vas", O.F1, O.g1_compress), ("wass", "was", O.F1, O.g1_compress),
                            ("yass", "yas", O.F1, O.g1_compress), ("wss", "ws", O.F2, O.g2_compress)):
        assert getattr(pp, f) == comp(O.msm_naive(F, fe, st["EK"][key])), f
    gz = O.g1_add(O.msm_naive(O.F1, fe, st["EK"]["vbs"]), O.g1_add(O.msm_naive(O.F1, fe, st["EK"]["wbs"]),
                                                                    O.msm_naive(O.F1, fe, st["EK"]["ybs"])))
    assert pp.gz == O.g1_compress(gz)
    if timings is not None:   # bench.py: end-to-end latency of the two provers on this config (host bytes in / out)
        import time
        ek = H.mirror_phgr13_ek(st["EK"])
        wb = b"".join(v.to_bytes(32, "big") for v in w)
        for name, fn in (("groth16_prove_ms", lambda: api.Groth16Prove(tr, q, wb, rr, ss, backend=be)),
                         ("phgr13_prove_ms", lambda: api.PHGR13Prove(ek, q, wb, backend=be))):
            fn()
            best = 1e9
            for _ in range(5):
                t0 = time.perf_counter(); fn(); best = min(best, time.perf_counter() - t0)
            timings[name] = best * 1e3
        timings.update(gates=n, variables=oq.nb_vars, mid=oq.nb_io)


def readme_flow_through_api(be):
    """README / r1cs.go:178-198: createR1CS -> ToQAP -> Groth16Prove, all through the host mirror's
    reference-shaped API, against the golden proof."""
    c = api.R1CS()
    c.NewInput("x"); c.NewOutput("out")
    c.NewVar("u"); c.NewVar("v"); c.NewVar("w")
    c.Mul("x", "x", "u"); c.Mul("u", "x", "v"); c.Add("v", "x", "w"); c.AddConst("w", 5, "out")
    assert c.vars == ["const", "x", "out", "u", "v", "w"] and c.nbIO() == 3
    sol = [0] * 6
    for name, val in (("const", 1), ("x", 3), ("out", 35), ("u", 9), ("v", 27), ("w", 30)):
        sol[c.IndexOf(name)] = val                       # createWitness, r1cs.go:67-76
    q = api.ToQAP(c)
    g = gold("readme_circuit")
    assert q.nbVars == g["nb_vars"] and q.nbIO == g["nb_io"] and q.nbGates == g["nb_gates"]
    h, (a, b, cc) = api.Quotient(q, sol, backend=be, return_abc=True)
    I = lambda xs: [int(x, 16) for x in xs]
    assert h == I(g["h"]) and a == I(g["a"]) and b == I(g["b"]) and cc == I(g["c"])
    k = g["groth16"]
    B = bytes.fromhex
    tr = api.Groth16Setup(Alpha=B(k["Alpha"]), Beta=B(k["Beta"]), Delta=B(k["Delta"]), Xi=[B(x) for x in k["Xi"]],
                          NioLP=[B(x) for x in k["NioLP"]], XiT=[B(x) for x in k["XiT"]], Beta2=B(k["Beta2"]),
                          Delta2=B(k["Delta2"]), Xi2=[B(x) for x in k["Xi2"]])
    pr = api.Groth16Prove(tr, q, sol, int(k["r"], 16), int(k["s"], 16), backend=be)
    assert (pr.A.hex(), pr.B.hex(), pr.C.hex()) == (k["A"], k["B"], k["C"])
    ek = api.PHGR13EvalKey(**{name: [bytes.fromhex(x) for x in v] for name, v in g["phgr13"]["ek"].items()})
    pp = api.PHGR13Prove(ek, q, sol, backend=be)
    for f in O.PHGR13_FIELDS:
        assert getattr(pp, f).hex() == g["phgr13"]["proof"][f], f
    import pytest
    with pytest.raises(KeyError, match="plouf"):         # r1cs.go:27
        c.IndexOf("nope")
    c5 = api.R1CS(); c5.NewInput("x"); c5.NewOutput("o"); c5.Mul("x", "x", "o"); c5.Mul("x", "x", "o"); c5.Mul("x", "x", "o")
    # three gates (not a power of two): accepted like the reference's ToQAP; h z = a b - c at a random point
    q5 = api.ToQAP(c5)
    assert q5.nbGates == 3
    sol5 = [0] * 3
    for name, val in (("const", 1), ("x", 3), ("o", 9)):
        sol5[c5.IndexOf(name)] = val
    h5, (a5, b5, cc5) = api.Quotient(q5, sol5, backend=be, return_abc=True)
    x0 = 0x1234567
    z5 = (x0 - 1) * (x0 - 2) * (x0 - 3) % O.R
    assert len(h5) == 2 and len(a5) == 3
    assert O.poly_eval(h5, x0) * z5 % O.R == (O.poly_eval(a5, x0) * O.poly_eval(b5, x0) - O.poly_eval(cc5, x0)) % O.R
    assert all(O.poly_eval(a5, j) * O.poly_eval(b5, j) % O.R == O.poly_eval(cc5, j) for j in (1, 2, 3))


def phgr13_sparse_exponent_check(be, log_n, seed, n=None):
    """PHGR13 prove on a sparse synthetic circuit of 2^log_n gates: all eight proof elements must equal
    their exponent-level recomputation from the toxic waste (pinocchio_test.go:23-196 at scale)."""
    n = n or (1 << log_n)
    sq, wit = H.sparse_circuit(n, seed, n // 2)
    ek, tw = H.sparse_phgr13_setup(be, sq, seed)
    pp = api.PHGR13Prove(ek, sq, wit, backend=be, want_h=True)
    want = H.sparse_phgr13_expected(sq, wit, tw)
    for f in O.PHGR13_FIELDS:
        assert getattr(pp, f) == want[f], f
    assert len(pp.h) == n - 1


def multi_groth16_case(lib, be, ndev, log_n, seed, devices=None, rank0_share=None, n=None):
    """ps_mg16_prove on `ndev` devices (one host call, key sharded, sparse QAP replicated) must return the proof
    of ps_g16_prove on one device bit for bit; a broken witness must raise "apocalypse" from whichever device
    owns the gate; a witness of the wrong length the sanityCheck error (qap.go:177-189)."""
    import pytest
    n = n or (1 << log_n)      # any n >= 2: the interpolation tree pads to the next power of two with dummy leaves
    sq, wit = H.sparse_circuit(n, seed, n // 2)
    tr, tw = H.sparse_groth16_setup(be, sq, seed)
    smp = O.Sampler(seed + 5)
    r, s = smp.fr(), smp.fr()
    want = api.Groth16Prove(tr, sq, wit, r, s, backend=be)
    A, B, Cc, _ = H.sparse_groth16_expected(sq, wit, tw, r, s)
    assert (want.A, want.B, want.C) == (A, B, Cc)
    sq.close(); tr.close()
    mb = api.MultiBackend(devices if devices is not None else [0] * ndev, lib=lib)
    try:
        if rank0_share is not None:
            mb.set_option("rank0_share_percent", rank0_share)
        for _ in range(2):      # second call: workspaces and tables are reused
            got = api.Groth16Prove(tr, sq, wit, r, s, backend=mb)
            assert (got.A, got.B, got.C) == (A, B, Cc)
        bad = list(wit); bad[-1] = (bad[-1] + 1) % O.R
        with pytest.raises(ArithmeticError, match="apocalypse"):
            api.Groth16Prove(tr, sq, bad, r, s, backend=mb)
        with pytest.raises(ValueError):
            api.Groth16Prove(tr, sq, wit[:-1], r, s, backend=mb)
        got = api.Groth16Prove(tr, sq, wit, r, s, backend=mb)       # still healthy after the failed calls
        assert (got.A, got.B, got.C) == (A, B, Cc)
        tl = mb.timeline(0)
    finally:
        sq.close(); tr.close()
        mb.close()
    return tl


def multi_groth16_dense_case(lib, be, ndev, devices=None):
    """dense QAP (configs[0] / [1] shapes): the quotient runs on device 0, the MSM shards everywhere"""
    c = O.create_r1cs(); w = O.create_witness(c); oq = O.to_qap(c)
    smp = O.Sampler(0)
    otr = O.groth16_setup(oq, smp)
    r, s = smp.fr(), smp.fr()
    want = O.groth16_prove(otr, oq, w, r, s)
    tr, q = H.mirror_g16_setup(otr), H.mirror_qap(oq)
    mb = api.MultiBackend(devices if devices is not None else [0] * ndev, lib=lib)
    try:
        got = api.Groth16Prove(tr, q, w, r, s, backend=mb)
        assert (got.A, got.B, got.C) == (O.g1_compress(want["A"]), O.g2_compress(want["B"]), O.g1_compress(want["C"]))
    finally:
        q.close(); tr.close()
        mb.close()


def multi_msm_case(lib, ndev, group, n, devices=None, window_bits=0, tables=1):
    import pytest
    F, gen, comp, _ = _grp(group)
    rng = random.Random(n * 31 + ndev)
    ks = [rng.randrange(1, O.R) for _ in range(n)]
    sc = [rng.randrange(O.R) for _ in range(n)]
    mb = api.MultiBackend(devices if devices is not None else [0] * ndev, lib=lib)
    try:
        bases = mb.bases_from_scalars(group, ks, window_bits, tables)
        assert len(bases) == n
        want = comp(O.pt_mul(F, sum(k * s for k, s in zip(ks, sc)) % O.R, gen))
        assert mb.msm(bases, sc) == want
        with pytest.raises(ValueError):
            mb.msm(bases, sc[:-1])
        with pytest.raises(api.L.PlaysnarkError):
            mb.msm(bases, b"\xff" * (32 * n))
        # the same points through the wire format
        single = api.Backend(0, lib=lib)
        pts = single.bases_from_scalars(group, ks).export()
        single.close()
        b2 = mb.load_bases(group, pts)
        assert mb.msm(b2, sc) == want
        b2.close(); bases.close()
    finally:
        mb.close()


def device_setups_vs_oracle(be, n=12, seed=3):
    """ps_g16_setup / ps_phgr13_setup (the trusted setups on the device, from injected toxic waste) against the
    oracle's NewGroth16TrustedSetup / NewPHGR13TrustedSetup restatements: every key element byte for byte, for the
    dense QAP and for the sparse R1CS form of the same circuit; then prove with the device-made keys."""
    r, w = H.mixed_circuit(n, seed, max(1, n // 2)) if n & (n - 1) == 0 else H.mixed_circuit(n, seed, max(1, n // 2))
    oq = O.to_qap(r)
    dense = H.mirror_qap(oq)
    forms = [dense, api.SparseQAP.from_dense_rows(len(r.vars), r.nb_io(), r.left, r.right, r.out)]
    g1, g2 = O.g1_compress, O.g2_compress
    otr = O.groth16_setup(oq, O.Sampler(seed))
    toxic = tuple(otr.tw[k] for k in ("Alpha", "Beta", "Delta", "X", "Gamma"))
    smp = O.Sampler(seed + 50)
    rr, ss = smp.fr(), smp.fr()
    want = O.groth16_prove(otr, oq, w, rr, ss)
    ost = O.phgr13_setup(oq, O.Sampler(seed + 1))
    s2 = O.Sampler(seed + 1)
    ptoxic = [s2.fr() for _ in range(7)]            # s av aw ay rv rw beta
    ptoxic.append(s2.fr())                          # gamma
    pwant = O.phgr13_prove(ost["EK"], oq, w)
    for q in forms:
        tr = api.NewGroth16TrustedSetup(q, backend=be, toxic=toxic, fmt=L.PS_FMT_COMPRESSED)
        cut = lambda raw, per: [raw[i:i + per] for i in range(0, len(raw), per)]
        assert cut(tr.Xi, 48) == [g1(p) for p in otr.Xi] and cut(tr.Xi2, 96) == [g2(p) for p in otr.Xi2]
        assert cut(tr.XiT, 48) == [g1(p) for p in otr.XiT] and cut(tr.NioLP, 48) == [g1(p) for p in otr.NioLP]
        assert tr.IoLP == [g1(p) for p in otr.IoLP] and tr.Gamma == g2(otr.Gamma)
        assert (tr.Alpha, tr.Beta, tr.Delta, tr.Beta2, tr.Delta2) == (g1(otr.Alpha), g1(otr.Beta), g1(otr.Delta), g2(otr.Beta2), g2(otr.Delta2))
        pr = api.Groth16Prove(tr, q, w, rr, ss, backend=be)       # resident key straight from the setup
        assert (pr.A, pr.B, pr.C) == (g1(want["A"]), g2(want["B"]), g1(want["C"]))
        tr.close()
        pr = api.Groth16Prove(tr, q, w, rr, ss, backend=be)       # the same key reloaded from its exported bytes
        assert (pr.A, pr.B, pr.C) == (g1(want["A"]), g2(want["B"]), g1(want["C"]))
        tr.close()
        ek, vk, _ = api.NewPHGR13TrustedSetup(q, backend=be, toxic=ptoxic, with_vk=True)
        pp = api.PHGR13Prove(ek, q, w, backend=be)
        for f in O.PHGR13_FIELDS:
            assert getattr(pp, f) == (g2 if f == "wss" else g1)(pwant[f]), f
        ek.export()
        for name, pts in ost["EK"].items():
            assert getattr(ek, name) == [(g2 if name == "ws" else g1)(p) for p in pts], name
        ovk = ost["VK"]
        assert vk["av"] == g2(ovk["av"]) and vk["aw"] == g1(ovk["aw"]) and vk["ay"] == g2(ovk["ay"])
        assert vk["gamma"] == g2(ovk["gamma"]) and vk["bgamma"] == g1(ovk["bgamma"]) and vk["bgamma2"] == g2(ovk["bgamma2"])
        assert vk["yts"] == g2(ovk["yts"])
        assert vk["vs"] == [g1(p) for p in ovk["vs"]] and vk["ws"] == [g2(p) for p in ovk["ws"]] and vk["ys"] == [g1(p) for p in ovk["ys"]]
        ek.close(); q.close()


# ---- verifiers: the device decides every case the way the oracle's verifier does ----------------
def pairing_checks(be):
    """bilinearity / non-degeneracy of the device pairing product (the only property the verifiers observe), infinity
    on either side, several checks in one call"""
    a, b = 0x1234567, 0x7654321
    g1 = lambda k: O.g1_compress(O.g1_mul(k % O.R))
    g2 = lambda k: O.g2_compress(O.g2_mul(k % O.R))
    inf1, inf2 = O.g1_compress(None), O.g2_compress(None)
    P1 = [g1(a), g1(-a * b), g1(a), g1(-a * b + 1), g1(1), inf1, g1(a), g1(5), g1(7), g1(-(5 * 11 + 7 * 13))]
    Q2 = [g2(b), g2(1), g2(b), g2(1), g2(1), g2(3), inf2, g2(11), g2(13), g2(1)]
    res = api.PairingCheckBatch(P1, Q2, [2, 2, 1, 2, 3], backend=be)
    #      e(aG,bH)e(-abG,H)=1 | off by one | e(G,H) != 1 | two infinities -> 1 | e(5G,11H)e(7G,13H)e(-(55+91)G,H)=1
    assert res == [True, False, False, True, True], res
    assert res == [O.fp12_mul(O.pairing(O.g1_decompress(P1[0]), O.g2_decompress(Q2[0])),
                              O.pairing(O.g1_decompress(P1[1]), O.g2_decompress(Q2[1]))) == O.FP12_ONE,
                   False, False, True, True]
    import pytest
    from playsnark_b200._lib import PlaysnarkError
    bad = bytearray(g1(a)); bad[-1] ^= 1                      # not on the curve
    with pytest.raises(PlaysnarkError):
        api.PairingCheckBatch([bytes(bad)], [g2(b)], [1], backend=be)


def groth16_verify_cases(be, n=8, seed=9):
    """Groth16Verify (groth16.go:214-233) on the device against the oracle's verifier: an honest proof (from the oracle's
    prover and from the device's), a mutated A / C, a wrong public input, a wrong Gamma"""
    r1, w = H.mixed_circuit(n, seed, max(1, n // 2))
    oq = O.to_qap(r1)
    smp = O.Sampler(seed)
    otr = O.groth16_setup(oq, smp)
    rr, ss = smp.fr(), smp.fr()
    op = O.groth16_prove(otr, oq, w, rr, ss)
    diff = oq.nb_vars - oq.nb_io
    io = w[:diff]
    assert O.groth16_verify(otr, oq, op, io)
    tr = H.mirror_g16_setup(otr)
    q = H.mirror_qap(oq)
    proof = api.Groth16Proof(tp=api.Groth16ToxicProof(rr, ss), A=O.g1_compress(op["A"]), B=O.g2_compress(op["B"]), C=O.g1_compress(op["C"]))
    assert api.Groth16Verify(tr, q, proof, io, backend=be)
    dev = api.Groth16Prove(tr, q, w, rr, ss, backend=be)                       # the device's own proof
    assert (dev.A, dev.B, dev.C) == (proof.A, proof.B, proof.C) and api.Groth16Verify(tr, q, dev, io, backend=be)
    for f in ("A", "C"):
        bad_o = dict(op); bad_o[f] = O.g1_add(op[f], O.G1_GEN)
        bad = api.Groth16Proof(tp=proof.tp, A=O.g1_compress(bad_o["A"]), B=proof.B, C=O.g1_compress(bad_o["C"]))
        assert not O.groth16_verify(otr, oq, bad_o, io)
        assert not api.Groth16Verify(tr, q, bad, io, backend=be), f
    if diff > 1:
        io2 = list(io); io2[-1] = (io2[-1] + 1) % O.R
        assert not O.groth16_verify(otr, oq, op, io2) and not api.Groth16Verify(tr, q, proof, io2, backend=be)
    tr2 = H.mirror_g16_setup(otr); tr2.Gamma = O.g2_compress(O.g2_add(otr.Gamma, O.G2_GEN))
    assert not api.Groth16Verify(tr2, q, proof, io, backend=be)
    tr.close(); q.close()


def phgr13_verify_cases(be, n=6, seed=4):
    """PHGR13Verify (pinochio.go:281-375) on the device: the honest proof verifies; the five mutated proof fields and
    the three mutated verification-key fields of pinocchio_test.go:243-276 are rejected -- the oracle's decisions"""
    r1, w = H.mixed_circuit(n, seed, max(1, n // 2))
    oq = O.to_qap(r1)
    st = O.phgr13_setup(oq, O.Sampler(seed + 1))
    pp = O.phgr13_prove(st["EK"], oq, w)
    diff = oq.nb_vars - oq.nb_io
    io = w[:diff]
    assert O.phgr13_verify(st["VK"], oq, pp, io)
    g1, g2 = O.g1_compress, O.g2_compress

    def mirror_vk(vk):
        return {"av": g2(vk["av"]), "aw": g1(vk["aw"]), "ay": g2(vk["ay"]), "gamma": g2(vk["gamma"]), "bgamma": g1(vk["bgamma"]),
                "bgamma2": g2(vk["bgamma2"]), "yts": g2(vk["yts"]), "vs": [g1(p) for p in vk["vs"]], "ws": [g2(p) for p in vk["ws"]],
                "ys": [g1(p) for p in vk["ys"]]}

    def mirror_proof(d):
        return api.PHGR13Proof(**{f: (g2 if f == "wss" else g1)(d[f]) for f in O.PHGR13_FIELDS})
    q = H.mirror_qap(oq)
    assert api.PHGR13Verify(mirror_vk(st["VK"]), q, mirror_proof(pp), io, backend=be)
    dev = api.PHGR13Prove(H.mirror_phgr13_ek(st["EK"]), q, w, backend=be)      # the device's own proof
    assert api.PHGR13Verify(mirror_vk(st["VK"]), q, dev, io, backend=be)
    for f in ("hs", "vss", "vass", "wass", "yass"):
        bad = dict(pp); bad[f] = O.g1_add(pp[f], O.G1_GEN)
        assert not O.phgr13_verify(st["VK"], oq, bad, io)
        assert not api.PHGR13Verify(mirror_vk(st["VK"]), q, mirror_proof(bad), io, backend=be), f
    for f in ("yts", "gamma", "bgamma2"):
        vk = dict(st["VK"]); vk[f] = O.g2_add(vk[f], O.G2_GEN)
        assert not O.phgr13_verify(vk, oq, pp, io)
        assert not api.PHGR13Verify(mirror_vk(vk), q, mirror_proof(pp), io, backend=be), f
    io2 = list(io); io2[0] = (io2[0] + 1) % O.R
    assert not api.PHGR13Verify(mirror_vk(st["VK"]), q, mirror_proof(pp), io2, backend=be)
    q.close()


def verify_device_setup_flow(be, n=16, seed=2):
    """setup -> prove -> verify entirely through the device entry points (TestGroth16Verify groth16_test.go:22-30 and the
    end of TestPinocchioProofValidDivision), sparse QAP"""
    sq, wit = H.sparse_circuit(n, seed, n // 2)
    diff = sq.nbVars - sq.nbIO
    tr = api.NewGroth16TrustedSetup(sq, backend=be, fmt=L.PS_FMT_COMPRESSED)
    pr = api.Groth16Prove(tr, sq, wit, 12345, 67890, backend=be)
    assert api.Groth16Verify(tr, sq, pr, wit[:diff], backend=be)
    bad = list(wit[:diff]); bad[0] = (bad[0] + 1) % O.R
    assert not api.Groth16Verify(tr, sq, pr, bad, backend=be)
    ek, vk, _ = api.NewPHGR13TrustedSetup(sq, backend=be, with_vk=True)
    pp = api.PHGR13Prove(ek, sq, wit, backend=be)
    assert api.PHGR13Verify(vk, sq, pp, wit[:diff], backend=be)
    assert not api.PHGR13Verify(vk, sq, pp, bad, backend=be)
    tr.close(); ek.close(); sq.close()

