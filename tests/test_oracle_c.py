"""The C part of the oracle (oracle/ps_oracle.c) against the Python oracle."""
import random

import pytest

from oracle import c_oracle as CO, ps_oracle as O
from tests import helpers as H


def test_field_ops():
    rng = random.Random(1)
    for _ in range(200):
        a, b = rng.randrange(O.P), rng.randrange(O.P)
        assert CO.fp_mul(a, b) == a * b % O.P
        x, y = rng.randrange(O.R), rng.randrange(1, O.R)
        assert CO.fr_mul(x, y) == x * y % O.R
    assert CO.fp_mul(O.P - 1, O.P - 1) == 1 and CO.fr_mul(O.R - 1, O.R - 1) == 1
    assert CO.fr_inv(7) * 7 % O.R == 1


def test_scalar_mul_and_blind_eval():
    rng = random.Random(2)
    for k in [0, 1, 2, O.R - 1, rng.randrange(O.R), rng.randrange(1 << 64)]:
        assert CO.g1_mul(O.G1_GEN, k) == O.g1_mul(k)
    assert CO.g1_mul(None, 5) is None
    n = 9
    pts = [O.g1_mul(rng.randrange(1, O.R)) for _ in range(n)]
    pts[4] = None; pts[6] = pts[5]
    sc = [rng.randrange(O.R) for _ in range(n)]
    sc[0] = 0
    got = CO.blind_eval_g1(b"".join(O.g1_affine_bytes(p) for p in pts), b"".join(O.fr_to_bytes(s) for s in sc))
    assert got == O.msm_naive(O.F1, sc, pts)


@pytest.mark.parametrize("n", [4, 7, 16])
def test_quotient_and_aggregate(n):
    if n == 4:
        c = O.create_r1cs(); w = O.create_witness(c)
    else:
        c, w = H.squaring_chain(n, 12345 + n)
    q = O.to_qap(c)
    a, b, cc = q.compute_aggregate_poly(w)
    assert CO.aggregate(q.left, [O.value_to_fr(v) for v in w]) == a
    for faithful in (False, True):
        assert CO.quotient(a, b, cc, q.z, faithful) == q.quotient(w)
    bad = list(cc); bad[0] = (bad[0] + 1) % O.R
    with pytest.raises(ArithmeticError):
        CO.quotient(a, b, bad, q.z)


def _toxic_g16(seed):
    smp = O.Sampler(seed)
    tox = [smp.fr() for _ in range(4)]           # alpha, beta, delta, x (NewGroth16TrustedSetup's order)
    return tox, smp


@pytest.mark.parametrize("case", ["readme", "mixed12", "chain8_neg"])
def test_groth16_flow_matches_python_oracle(case):
    """ToQAP -> setup -> Groth16Prove restated in C (oracle/ps_prover.c) against the Python oracle: the
    same A, B (G2), C affine coordinates and h, single- and multi-threaded."""
    if case == "readme":
        c = O.create_r1cs(); w = O.create_witness(c)
    elif case == "mixed12":
        c, w = H.mixed_circuit(12, 5, 6)
    else:
        c, w = H.squaring_chain(8, O.R - 1)
    q = O.to_qap(c, fast=False)
    tox, smp = _toxic_g16(11)
    tr = O.groth16_setup(q, O.Sampler(11))
    assert [tr.tw[k] for k in ("Alpha", "Beta", "Delta", "X")] == tox
    smp.fr()                                      # gamma
    r, s = smp.fr(), smp.fr()
    want = O.groth16_prove(tr, q, w, r, s)
    for threads, fast in ((1, False), (3, False), (2, True)):
        A, B, C, h, sec = CO.groth16_flow(c, [O.value_to_fr(v) for v in w], tox, r, s, threads, fast_qap=fast)
        assert (A, B, C) == (want["A"], want["B"], want["C"])
        assert h == want["h"]
        assert all(v >= 0 for v in sec.values())
    bad = list(w); bad[-1] = (bad[-1] + 1) % O.R
    with pytest.raises(ArithmeticError):
        CO.groth16_flow(c, [O.value_to_fr(v) for v in bad], tox, r, s, 2)


def test_phgr13_flow_matches_python_oracle():
    c, w = H.mixed_circuit(10, 9, 5)
    q = O.to_qap(c, fast=False)
    st = O.phgr13_setup(q, O.Sampler(4), with_vk=False)
    smp = O.Sampler(4)
    s, av, aw, ay, rv, rw, beta = (smp.fr() for _ in range(7))   # pinochio.go:93-141 sampling order
    assert (st["t"]["s"], st["t"]["rv"], st["t"]["rw"], st["t"]["beta"]) == (s, rv, rw, beta)
    want = O.phgr13_prove(st["EK"], q, w)
    proof, h, _ = CO.phgr13_flow(c, [O.value_to_fr(v) for v in w], (s, av, aw, ay, rv, rw, beta), threads=2)
    for f in O.PHGR13_FIELDS:
        assert proof[f] == want[f], f
    assert h == want["h"]
    proof2, h2, _ = CO.phgr13_flow(c, [O.value_to_fr(v) for v in w], (s, av, aw, ay, rv, rw, beta), threads=1, fast_qap=True)
    assert proof2 == proof and h2 == h


def test_blind_eval_threads():
    rng = random.Random(8)
    n = 37
    pts = [O.g1_mul(rng.randrange(1, O.R)) for _ in range(n)]
    pts[3] = None
    sc = [rng.randrange(O.R) for _ in range(n)]
    pb = b"".join(O.g1_affine_bytes(p) for p in pts); sb = b"".join(O.fr_to_bytes(s) for s in sc)
    assert CO.blind_eval_g1_mt(pb, sb, 4) == CO.blind_eval_g1(pb, sb) == O.msm_naive(O.F1, sc, pts)
