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
