"""Device field / curve templates, compiled for the host with an emulated carry flag, against the
oracle.  Catches arithmetic mistakes without a GPU; the same templates are what the kernels run."""
import ctypes
import random

import pytest

from oracle import ps_oracle as O

U32 = ctypes.c_uint32


def limbs(x, n):
    return (U32 * n)(*[(x >> (32 * i)) & 0xFFFFFFFF for i in range(n)])


def unl(a):
    return sum(int(v) << (32 * i) for i, v in enumerate(a))


def fp_vals(rng, k):
    edge = [0, 1, 2, O.P - 1, O.P - 2, (1 << 384) % O.P, (1 << 380), (O.P - 1) // 2, 0xFFFFFFFF, 1 << 32]
    return edge + [rng.randrange(O.P) for _ in range(k)]


def fr_vals(rng, k):
    edge = [0, 1, 2, O.R - 1, O.R - 2, (1 << 256) % O.R, 1 << 254, 0xFFFFFFFF, 1 << 32]
    return edge + [rng.randrange(O.R) for _ in range(k)]


def test_fp_ops(host_check):
    rng = random.Random(1)
    vals = fp_vals(rng, 40)
    out = (U32 * 12)()
    for a in vals:
        for b in vals[:14] + vals[-6:]:
            host_check.hc_fp_mul(limbs(a, 12), limbs(b, 12), out); assert unl(out) == a * b % O.P
            host_check.hc_fp_add(limbs(a, 12), limbs(b, 12), out); assert unl(out) == (a + b) % O.P
            host_check.hc_fp_sub(limbs(a, 12), limbs(b, 12), out); assert unl(out) == (a - b) % O.P
    rinv = pow(1 << 384, -1, O.P)
    for a in vals:
        for b in vals[:12]:
            host_check.hc_fp_montmul_raw(limbs(a, 12), limbs(b, 12), out)
            assert unl(out) == a * b * rinv % O.P
    for a in vals[1:20]:
        host_check.hc_fp_inv(limbs(a, 12), out); assert unl(out) * a % O.P == 1
    host_check.hc_fp_inv(limbs(0, 12), out); assert unl(out) == 0
    for a in vals[:20]:
        sq = a * a % O.P
        host_check.hc_fp_sqrt(limbs(sq, 12), out); assert unl(out) in (a, (O.P - a) % O.P)


def test_inversion_binary_gcd_vs_fermat(host_check):
    """fp_inv_serial / fr_inv_serial (binary extended Euclid) against a^-1 mod p and against the Fermat route, edge values
    included (1, p-1, powers of two, values whose Montgomery form is small)."""
    rng = random.Random(9)
    out, out2 = (U32 * 12)(), (U32 * 12)()
    rinv = pow(1 << 384, -1, O.P)
    vals = fp_vals(rng, 120) + [rinv, 2 * rinv % O.P, (O.P - rinv) % O.P, 1 << 200, (1 << 381) % O.P, O.P - 1]
    for a in vals:
        host_check.hc_fp_inv_serial(limbs(a, 12), out)
        host_check.hc_fp_inv(limbs(a, 12), out2)
        assert unl(out) == unl(out2) == (pow(a, -1, O.P) if a else 0), hex(a)
    out, out2 = (U32 * 8)(), (U32 * 8)()
    rinv = pow(1 << 256, -1, O.R)
    vals = fr_vals(rng, 120) + [rinv, 2 * rinv % O.R, (O.R - rinv) % O.R, 1 << 200, O.R - 1]
    for a in vals:
        host_check.hc_fr_inv_serial(limbs(a, 8), out)
        host_check.hc_fr_inv(limbs(a, 8), out2)
        assert unl(out) == unl(out2) == (pow(a, -1, O.R) if a else 0), hex(a)


def test_to_affine_serial_matches(host_check):
    rng = random.Random(4)
    out = (U32 * 96)()
    for pre in (0, 1, 3):
        p = O.pt_mul(O.F2, rng.randrange(1, O.R), O.G2_GEN)
        arr = (U32 * 48)(*(list(limbs(p[0][0], 12)) + list(limbs(p[0][1], 12)) + list(limbs(p[1][0], 12)) + list(limbs(p[1][1], 12))))
        host_check.hc_g2_to_affine_serial(arr, pre, out)
        assert list(out[:48]) == list(out[48:])


def test_squaring_edge_values(host_check):
    """Fe::sqr against a*a mod p, on edge values (carry-heavy limbs) and random ones, through the domain conversion
    and on raw limbs."""
    rng = random.Random(7)
    out12, out8 = (U32 * 12)(), (U32 * 8)()
    heavy_p = [O.P - 1, O.P - 2, (1 << 380) - 1, (1 << 381) - 1 - ((1 << 381) - 1 >= O.P) * (1 << 380), int("f" * 95, 16) % O.P,
               sum(0xFFFFFFFF << (64 * k) for k in range(6)) % O.P, sum(0xFFFFFFFF << (64 * k + 32) for k in range(6)) % O.P]
    rinv = pow(1 << 384, -1, O.P)
    for a in fp_vals(rng, 300) + heavy_p:
        host_check.hc_fp_sqr(limbs(a, 12), out12); assert unl(out12) == a * a % O.P, hex(a)
        host_check.hc_fp_montsqr_raw(limbs(a, 12), out12); assert unl(out12) == a * a * rinv % O.P, hex(a)
    heavy_r = [O.R - 1, O.R - 2, (1 << 254) - 1, (1 << 255) - 1 - O.R if (1 << 255) - 1 >= O.R else (1 << 255) - 1 - (1 << 254),
               sum(0xFFFFFFFF << (64 * k) for k in range(4)) % O.R, sum(0xFFFFFFFF << (64 * k + 32) for k in range(4)) % O.R]
    rinv = pow(1 << 256, -1, O.R)
    for a in fr_vals(rng, 300) + heavy_r:
        host_check.hc_fr_sqr(limbs(a, 8), out8); assert unl(out8) == a * a % O.R, hex(a)
        host_check.hc_fr_montsqr_raw(limbs(a, 8), out8); assert unl(out8) == a * a * rinv % O.R, hex(a)


def test_fr_ops(host_check):
    rng = random.Random(2)
    vals = fr_vals(rng, 40)
    out = (U32 * 8)()
    rinv = pow(1 << 256, -1, O.R)
    for a in vals:
        for b in vals[:14] + vals[-6:]:
            host_check.hc_fr_mul(limbs(a, 8), limbs(b, 8), out); assert unl(out) == a * b % O.R
            host_check.hc_fr_add(limbs(a, 8), limbs(b, 8), out); assert unl(out) == (a + b) % O.R
            host_check.hc_fr_sub(limbs(a, 8), limbs(b, 8), out); assert unl(out) == (a - b) % O.R
            host_check.hc_fr_montmul_raw(limbs(a, 8), limbs(b, 8), out); assert unl(out) == a * b * rinv % O.R
    for a in vals[1:20]:
        host_check.hc_fr_inv(limbs(a, 8), out); assert unl(out) * a % O.R == 1


def f2l(a):
    return limbs(a[0] | (a[1] << 384), 24)


def unf2(a):
    v = unl(a)
    return (v & ((1 << 384) - 1), v >> 384)


def test_fp2_ops(host_check):
    rng = random.Random(3)
    vals = [(0, 0), (1, 0), (0, 1), (O.P - 1, O.P - 1)] + [(rng.randrange(O.P), rng.randrange(O.P)) for _ in range(20)]
    out = (U32 * 24)()
    for a in vals:
        for b in vals:
            host_check.hc_fp2_mul(f2l(a), f2l(b), out); assert unf2(out) == O.F2.mul(a, b)
        host_check.hc_fp2_sqr(f2l(a), out); assert unf2(out) == O.F2.sqr(a)
        if a != (0, 0):
            host_check.hc_fp2_inv(f2l(a), out); assert O.F2.mul(unf2(out), a) == (1, 0)


def test_unreduced_products(host_check):
    """Building blocks of mul_sub_pair (curve.cuh: Y3 = a b - c d with ONE Montgomery reduction) on raw limbs: the full
    768-bit product (operands up to 2^384 - 1, not only field elements), the reduction of any T < p R, and
    (a b + (p - c) d) / R against a b - c d in the Montgomery domain, edge values included."""
    rng = random.Random(5)
    P = O.P
    wide = (U32 * 24)()
    full = (1 << 384) - 1
    ints = [0, 1, full, full - 1, 1 << 383, 2 * P - 2, P, 0xFFFFFFFF, (1 << 352) - 1] + [rng.randrange(1 << 384) for _ in range(30)]
    for x in ints:
        for y in ints:
            host_check.hc_fp_mul_wide_raw(limbs(x, 12), limbs(y, 12), wide)
            assert unl(wide) == x * y
    # a b + c d unreduced, limbs drawn from saturating values: every carry out of a row's chain lands in a limb that
    # the FIRST product has already filled (a carry lost there shows up about once per 2^32 random rows: one wrong
    # point per 2^24-point MSM, invisible to random operands)
    sat = [0, 1, 0xFFFFFFFF, 0xFFFFFFFE, 0x80000000, 0x7FFFFFFF, 2]
    def sat_int(top):
        v = 0
        for i in range(12):
            v |= rng.choice(sat) << (32 * i)
        return v % top
    for _ in range(20000):
        a, b = sat_int(1 << 383), sat_int(1 << 384)
        c, d = sat_int(1 << 383), sat_int(1 << 384)
        host_check.hc_fp_mul2_wide_raw(limbs(a, 12), limbs(b, 12), limbs(c, 12), limbs(d, 12), wide)
        assert unl(wide) == a * b + c * d, (hex(a), hex(b), hex(c), hex(d))
        host_check.hc_fp_mul_wide_raw(limbs(b, 12), limbs(d, 12), wide)
        assert unl(wide) == b * d
    rinv = pow(1 << 384, -1, P)
    red = (U32 * 12)()
    for _ in range(3000):
        t = (sat_int(P) << 384) | sat_int(1 << 384)                  # any T < p R with saturated limbs
        host_check.hc_fp_redc_wide_raw(limbs(t, 24), red)
        assert unl(red) == t * rinv % P, hex(t)
        a, b, c, d = sat_int(P), sat_int(P), sat_int(P), sat_int(P)
        host_check.hc_fp_mul2_lazy_raw(limbs(a, 12), limbs(b, 12), limbs(c, 12), limbs(d, 12), red)
        assert unl(red) == (a * b - c * d) * rinv % P
        host_check.hc_fp_montmul_raw(limbs(a, 12), limbs(b, 12), red)   # the plain product under the same operands
        assert unl(red) == a * b * rinv % P
    fe = [0, 1, P - 1, P - 2, (P - 1) // 2, (1 << 380)] + [rng.randrange(P) for _ in range(12)]
    for a in fe[:8]:
        for b in fe[:8]:
            for c in fe[:6] + fe[-3:]:
                for d in fe[:4] + fe[-3:]:
                    host_check.hc_fp_mul2_lazy_raw(limbs(a, 12), limbs(b, 12), limbs(c, 12), limbs(d, 12), red)
                    assert unl(red) == (a * b - c * d) * rinv % P
    ts = [0, 1, P, (1 << 384) - 1, 1 << 384, P << 384, (P << 384) - 1, (P - 1) << 384, ((P - 1) << 384) | ((1 << 384) - 1), 2 * P * P - 1]
    ts += [rng.randrange(P << 384) for _ in range(300)]
    for t in ts:
        if t >= P << 384:
            continue
        host_check.hc_fp_redc_wide_raw(limbs(t, 24), red)
        assert unl(red) == t * rinv % P, hex(t)


def g1l(p):
    return limbs(0 if p is None else p[0] | (p[1] << 384), 24)


def ung1(a):
    v = unl(a)
    x, y = v & ((1 << 384) - 1), v >> 384
    return None if (x, y) == (0, 0) else (x, y)


def g2l(p):
    if p is None:
        return limbs(0, 48)
    (x0, x1), (y0, y1) = p
    return limbs(x0 | (x1 << 384) | (y0 << 768) | (y1 << 1152), 48)


def ung2(a):
    v = unl(a)
    m = (1 << 384) - 1
    c = [(v >> (384 * i)) & m for i in range(4)]
    return None if c == [0, 0, 0, 0] else ((c[0], c[1]), (c[2], c[3]))


@pytest.mark.parametrize("grp", ["g1", "g2"])
def test_group_law(host_check, grp):
    rng = random.Random(4)
    F = O.F1 if grp == "g1" else O.F2
    gen = O.G1_GEN if grp == "g1" else O.G2_GEN
    enc, dec = (g1l, ung1) if grp == "g1" else (g2l, ung2)
    out = (U32 * (24 if grp == "g1" else 48))()
    madd = getattr(host_check, "hc_%s_madd" % grp)
    add = getattr(host_check, "hc_%s_add" % grp)
    mul = getattr(host_check, "hc_%s_mul" % grp)
    ks = [1, 2, 3, 5, O.R - 1, O.R - 2] + [rng.randrange(1, O.R) for _ in range(4)]
    pts = [None] + [O.pt_mul(F, k, gen) for k in ks]
    for p in pts:
        for q in pts:
            for pre in (1, 2, 3):
                want = O.pt_add(F, O.pt_mul(F, pre, p), q)
                madd(enc(p), enc(q), pre, out); assert dec(out) == want, (grp, "madd", pre)
            for pre in (1, 2):
                want = O.pt_add(F, O.pt_mul(F, pre, p), O.pt_mul(F, pre, q))
                add(enc(p), enc(q), pre, out); assert dec(out) == want, (grp, "add", pre)
    for k in [0, 1, 2, O.R - 1, O.R, rng.randrange(O.R), rng.randrange(1 << 64)]:
        for p in pts[:4]:
            mul(enc(p), limbs(k, 8), out); assert dec(out) == O.pt_mul(F, k, p)
