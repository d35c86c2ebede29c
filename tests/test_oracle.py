"""Pins the oracle: public constants, the reference's integer-level known answers, algebraic
self-checks of its own tests, and the committed golden fixtures (tools/make_golden.py)."""
import random
from fractions import Fraction

from oracle import ps_oracle as O
from tests.parity_cases import gold


def test_public_constants():
    c = gold("constants")
    assert O.on_curve(O.F1, O.G1_GEN) and O.on_curve(O.F2, O.G2_GEN)
    assert O.pt_mul(O.F1, O.R - 1, O.G1_GEN) == O.pt_neg(O.F1, O.G1_GEN)      # r*G = O
    assert O.pt_add(O.F2, O.pt_mul(O.F2, O.R - 1, O.G2_GEN), O.G2_GEN) is None
    assert O.g1_compress(O.G1_GEN).hex() == c["g1_generator_compressed"]
    assert c["g1_generator_compressed"].startswith("97f1d3a73197d7942695638c4fa9ac0f")
    assert O.g2_compress(O.G2_GEN).hex() == c["g2_generator_compressed"]
    assert c["g2_generator_compressed"].startswith("93e02b6052719f607dacd3a088274f65")
    # 2*G1: the widely published compressed encoding
    assert O.g1_compress(O.g1_mul(2)).hex() == c["g1_two_g"]
    assert c["g1_two_g"].startswith("a572cbea904d67468808c8eb50a9450c")
    w = O.fr_root_of_unity(32)
    assert pow(w, 1 << 31, O.R) == O.R - 1 and "%064x" % w == c["root_of_unity_2_32"]


def test_compress_roundtrip():
    rng = random.Random(1)
    for _ in range(6):
        k = rng.randrange(1, O.R)
        p, q = O.g1_mul(k), O.g2_mul(k)
        assert O.g1_decompress(O.g1_compress(p)) == p and O.g2_decompress(O.g2_compress(q)) == q
        assert O.g1_from_affine_bytes(O.g1_affine_bytes(p)) == p and O.g2_from_affine_bytes(O.g2_affine_bytes(q)) == q
    assert O.g1_decompress(O.g1_compress(None)) is None and O.g2_decompress(O.g2_compress(None)) is None


def test_reference_integer_kats():
    # TestAlgebraPolyMul, algebra_test.go:76-104: (1+2x)(3+x^2) = 3+6x+x^2+2x^3
    assert O.poly_mul([1, 2], [3, 0, 1]) == [3, 6, 1, 2]
    # TestAlgebraEval, algebra_test.go:10-19: 1+x at 1 = 2
    assert O.poly_eval([1, 1], 1) == 2
    # TestAlgebraPolyDivManual, algebra_test.go:48-74: (2x^3-6x^2+4x) / ((x-1)(x-2)), remainder 0, high degree first
    qv, rem = O.poly_div_synthetic([2, (-6) % O.R, 4, 0], [1, (-3) % O.R, 2])
    assert qv == [2, 0] and all(v == 0 for v in rem)
    # same division through Div2 (low degree first)
    q2, r2 = O.poly_div2([0, 4, (-6) % O.R, 2], [2, (-3) % O.R, 1])
    assert q2 == [0, 2] and O.poly_normalize(r2) == []
    # TestAlgebraPolyDiv, algebra_test.go:178-195: Div2 round trip on random polynomials
    rng = random.Random(2)
    for _ in range(5):
        a = [rng.randrange(O.R) for _ in range(7)]
        b = [rng.randrange(O.R) for _ in range(3)] + [rng.randrange(1, O.R)]
        qq, rr = O.poly_div2(a, b)
        assert O.poly_normalize(O.poly_add(O.poly_mul(qq, b), rr)) == O.poly_normalize(a)
    # TestAlgebraInterpolate, algebra_test.go:37-45
    ys = [rng.randrange(O.R) for _ in range(6)]
    p = O.interpolate(ys)
    assert [O.poly_eval(p, i + 1) for i in range(6)] == ys
    # TestMatrixTranspose, algebra_test.go:136-153
    assert O.mat_transpose([[1, 2, 3], [4, 5, 6]]) == [[1, 4], [2, 5], [3, 6]]


def test_readme_circuit_kats():
    c = O.create_r1cs()
    w = O.create_witness(c)
    assert w == [1, 3, 35, 9, 27, 30]
    # TestR1CSEquation, r1cs_test.go:10-30
    L_, R_, O_ = (O.mat_mul_vec(m, w) for m in (c.left, c.right, c.out))
    assert [a * b - o for a, b, o in zip(L_, R_, O_)] == [0, 0, 0, 0]
    for fast in (False, True):
        q = O.to_qap(c, fast=fast)
        # TestQAPManual, qap_test.go:28-36
        assert [O.poly_eval(q.left[1], i) for i in range(1, 5)] == [1, 0, 1, 0]
        # per-gate identity, qap_test.go:42-61
        a, b, cc = q.compute_aggregate_poly(w)
        for gate in range(1, 5):
            assert (O.poly_eval(a, gate) * O.poly_eval(b, gate) - O.poly_eval(cc, gate)) % O.R == 0
        assert q.is_valid(w)
        h = q.quotient(w)
        assert len(h) - 1 == len(q.z) - 1 - 2 == q.nb_gates - 2       # groth16_test.go:16-19
        fr = lambda x: x.numerator * pow(x.denominator, -1, O.R) % O.R
        assert h == [fr(Fraction(-11, 3)), fr(Fraction(307, 18)), fr(Fraction(-31, 9))]   # closed form
        assert a == [43, fr(Fraction(-220, 3)), fr(Fraction(77, 2)), fr(Fraction(-31, 6))]
        assert q.z == [24, (-50) % O.R, 35, (-10) % O.R, 1]
    bad = list(w); bad[2] = 36
    assert not q.is_valid(bad)
    g = gold("readme_circuit")
    assert ["%064x" % v for v in h] == g["h"]


def test_pairing_bilinear():
    e = O.pairing(O.G1_GEN, O.G2_GEN)
    assert e != O.FP12_ONE and O.fp12_pow(e, O.R) == O.FP12_ONE
    a, b = 0x1234567, 0xabcdef0123
    assert O.pairing(O.g1_mul(a), O.g2_mul(b)) == O.fp12_pow(e, a * b % O.R)
    assert O.fp12_mul(O.pairing(O.g1_mul(a), O.G2_GEN), O.pairing(O.g1_mul(b), O.G2_GEN)) == O.pairing(O.g1_mul(a + b), O.G2_GEN)


def test_blind_eval_is_evaluation():
    # TestPinocchioCombine, pinocchio_test.go:11-21
    rng = random.Random(3)
    p = [rng.randrange(O.R) for _ in range(5)]
    x = rng.randrange(O.R)
    pts = O.generate_powers_commit(O.F1, O.G1_GEN, x, 1, 4)
    assert O.blind_eval(O.F1, p, pts) == O.g1_mul(O.poly_eval(p, x))
    assert O.msm_fast(O.F1, p, pts) == O.msm_naive(O.F1, p, pts)


def test_groth16_golden_and_selfchecks():
    g = gold("readme_circuit")
    c = O.create_r1cs(); w = O.create_witness(c); q = O.to_qap(c)
    smp = O.Sampler(0)
    tr = O.groth16_setup(q, smp)
    r, s = smp.fr(), smp.fr()
    assert "%064x" % r == g["groth16"]["r"]
    fast = O.groth16_prove(tr, q, w, r, s)
    slow = O.groth16_prove(tr, q, w, r, s, faithful=True)      # sumBlind walked variable by variable
    assert fast == slow
    assert O.g1_compress(fast["A"]).hex() == g["groth16"]["A"]
    assert O.g2_compress(fast["B"]).hex() == g["groth16"]["B"]
    assert O.g1_compress(fast["C"]).hex() == g["groth16"]["C"]
    assert O.groth16_expected_from_toxic(tr, q, w, r, s) == (fast["A"], fast["B"], fast["C"])  # groth16_test.go:32-107
    diff = q.nb_vars - q.nb_io
    assert O.groth16_verify(tr, q, fast, w[:diff])                                              # groth16_test.go:22-30
    bad = dict(fast); bad["A"] = O.g1_add(fast["A"], O.G1_GEN)
    assert not O.groth16_verify(tr, q, bad, w[:diff])


def test_phgr13_golden_and_mutations():
    g = gold("readme_circuit")
    c = O.create_r1cs(); w = O.create_witness(c); q = O.to_qap(c)
    st = O.phgr13_setup(q, O.Sampler(1))
    pp = O.phgr13_prove(st["EK"], q, w)
    for f in O.PHGR13_FIELDS:
        enc = O.g2_compress if f == "wss" else O.g1_compress
        assert enc(pp[f]).hex() == g["phgr13"]["proof"][f]
    diff = q.nb_vars - q.nb_io
    io = w[:diff]
    assert O.phgr13_verify(st["VK"], q, pp, io)
    # hs == h(s) * G (pinocchio_test.go:31-44)
    assert pp["hs"] == O.g1_mul(O.poly_eval(pp["h"], st["t"]["s"]))
    # mutated proof fields and VK fields must fail (pinocchio_test.go:243-276)
    for f in ("hs", "vss", "vass", "wass", "yass"):
        bad = dict(pp); bad[f] = O.g1_add(pp[f], O.G1_GEN)
        assert not O.phgr13_verify(st["VK"], q, bad, io), f
    for f in ("yts", "gamma", "bgamma2"):
        vk = dict(st["VK"]); vk[f] = O.g2_add(vk[f], O.G2_GEN)
        assert not O.phgr13_verify(vk, q, pp, io), f
