"""CPU oracle for the playsnark proving path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product path (playsnark_b200/) never does and fails loudly without its CUDA library.

PARITY UNPINNED: the reference (nikkolasg/playsnark, Go) cannot be built here (no Go toolchain, and
its arithmetic lives in un-vendored modules: github.com/drand/kyber v1.1.3,
github.com/drand/kyber-bls12381 v0.2.1-0.20200920171356-02a6d1c7cc77, github.com/kilic/bls12-381
v0.0.0-20200820230200-6b2c19996391 -- go.mod:6-8), and none of its tests holds a hard-coded field
element, coordinate or serialised point.  This file is therefore a plain
big-integer restatement of the published algorithms (BLS12-381 short-Weierstrass group law, zcash
point encoding, optimal-ate pairing) anchored on
  * the public curve constants (p, r, generators, their compressed encodings),
  * the reference's integer-level known answers (algebra_test.go:10-19,48-74,76-104;
    qap_test.go:28-61; groth16_test.go:16-19) and the README circuit's closed-form quotient,
  * the algebraic self-checks the reference's own tests use (groth16_test.go:32-107,
    pinocchio_test.go:11-278): exponent-level recomputation, verifier acceptance, mutation rejects.
Group elements have canonical affine coordinates, so any correct implementation yields the same
bytes; the checks above are what pins "correct".

Every function cites the reference file:line it restates.
"""
from __future__ import annotations

import hashlib
from typing import List, Optional, Sequence, Tuple

# --------------------------------------------------------------------------------------------
# Constants (public BLS12-381 parameters; curve.go:13 selects this suite)
# --------------------------------------------------------------------------------------------
P = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
BLS_X = 0xd201000000010000  # |x|, the curve parameter is -x
G1_GEN = (
    0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb,
    0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1,
)
G2_GEN = (
    (0x024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8,
     0x13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e),
    (0x0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801,
     0x0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be),
)
FR_TWO_ADICITY = 32
FR_GENERATOR = 7


def fr_root_of_unity(log_n: int) -> int:
    """Primitive 2^log_n-th root of unity in Fr (7 is a generator of Fr^*)."""
    assert 0 <= log_n <= FR_TWO_ADICITY
    return pow(FR_GENERATOR, (R - 1) >> log_n, R)


# --------------------------------------------------------------------------------------------
# Field policies: F1 = Fp (ints), F2 = Fp2 = Fp[u]/(u^2+1) (pairs)
# --------------------------------------------------------------------------------------------
class F1:
    zero = 0
    one = 1
    b = 4  # y^2 = x^3 + 4

    @staticmethod
    def add(a, b): return (a + b) % P
    @staticmethod
    def sub(a, b): return (a - b) % P
    @staticmethod
    def neg(a): return (-a) % P
    @staticmethod
    def mul(a, b): return a * b % P
    @staticmethod
    def sqr(a): return a * a % P
    @staticmethod
    def inv(a): return pow(a, -1, P)
    @staticmethod
    def is_zero(a): return a == 0
    @staticmethod
    def muli(a, k): return a * k % P


class F2:
    zero = (0, 0)
    one = (1, 0)
    b = (4, 4)  # y^2 = x^3 + 4(1+u)

    @staticmethod
    def add(a, b): return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)
    @staticmethod
    def sub(a, b): return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)
    @staticmethod
    def neg(a): return ((-a[0]) % P, (-a[1]) % P)
    @staticmethod
    def mul(a, b):
        return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)
    @staticmethod
    def sqr(a):
        return ((a[0] + a[1]) * (a[0] - a[1]) % P, 2 * a[0] * a[1] % P)
    @staticmethod
    def inv(a):
        d = pow(a[0] * a[0] + a[1] * a[1], -1, P)
        return (a[0] * d % P, (-a[1]) * d % P)
    @staticmethod
    def is_zero(a): return a[0] == 0 and a[1] == 0
    @staticmethod
    def muli(a, k): return (a[0] * k % P, a[1] * k % P)


# --------------------------------------------------------------------------------------------
# Group law.  Affine points are (x, y) or None (infinity).  Jacobian (X, Y, Z), Z == 0 <=> infinity.
# kyber.Point.Add/Mul/Neg/Null/Base as called at algebra.go:356,365,373,381; groth16.go:138-200.
# --------------------------------------------------------------------------------------------
def on_curve(F, pt) -> bool:
    if pt is None:
        return True
    x, y = pt
    return F.sqr(y) == F.add(F.mul(F.sqr(x), x), F.b)


def to_jac(F, pt):
    return (F.one, F.one, F.zero) if pt is None else (pt[0], pt[1], F.one)


def to_affine(F, J):
    X, Y, Z = J
    if F.is_zero(Z):
        return None
    zi = F.inv(Z)
    zi2 = F.sqr(zi)
    return (F.mul(X, zi2), F.mul(Y, F.mul(zi2, zi)))


def jac_dbl(F, J):
    X, Y, Z = J
    if F.is_zero(Z) or F.is_zero(Y):
        return (F.one, F.one, F.zero)
    A = F.sqr(X); B = F.sqr(Y); C = F.sqr(B)
    D = F.muli(F.sub(F.sqr(F.add(X, B)), F.add(A, C)), 2)
    E = F.muli(A, 3); Fq = F.sqr(E)
    X3 = F.sub(Fq, F.muli(D, 2))
    Y3 = F.sub(F.mul(E, F.sub(D, X3)), F.muli(C, 8))
    Z3 = F.muli(F.mul(Y, Z), 2)
    return (X3, Y3, Z3)


def jac_add(F, J1, J2):
    X1, Y1, Z1 = J1
    X2, Y2, Z2 = J2
    if F.is_zero(Z1):
        return J2
    if F.is_zero(Z2):
        return J1
    Z1Z1 = F.sqr(Z1); Z2Z2 = F.sqr(Z2)
    U1 = F.mul(X1, Z2Z2); U2 = F.mul(X2, Z1Z1)
    S1 = F.mul(Y1, F.mul(Z2, Z2Z2)); S2 = F.mul(Y2, F.mul(Z1, Z1Z1))
    if U1 == U2:
        if S1 == S2:
            return jac_dbl(F, J1)
        return (F.one, F.one, F.zero)
    H = F.sub(U2, U1); Rr = F.sub(S2, S1)
    HH = F.sqr(H); HHH = F.mul(H, HH); V = F.mul(U1, HH)
    X3 = F.sub(F.sub(F.sqr(Rr), HHH), F.muli(V, 2))
    Y3 = F.sub(F.mul(Rr, F.sub(V, X3)), F.mul(S1, HHH))
    Z3 = F.mul(F.mul(Z1, Z2), H)
    return (X3, Y3, Z3)


def pt_add(F, a, b):
    return to_affine(F, jac_add(F, to_jac(F, a), to_jac(F, b)))


def pt_neg(F, a):
    return None if a is None else (a[0], F.neg(a[1]))


def pt_mul(F, k: int, a):
    """Scalar multiplication k*a, k reduced mod r (kyber Scalar is an element of Fr); k = 0 or
    a = infinity give infinity, as the reference's bit-serial double-and-add does."""
    k %= R
    acc = (F.one, F.one, F.zero)
    if a is None or k == 0:
        return None
    base = to_jac(F, a)
    for bit in bin(k)[2:]:
        acc = jac_dbl(F, acc)
        if bit == "1":
            acc = jac_add(F, acc, base)
    return to_affine(F, acc)


def in_subgroup(F, a) -> bool:
    """r * a == O with the scalar NOT reduced (kilic's FromCompressed / FromBytes reject points of the
    curve that lie outside the prime-order subgroup; kyber's UnmarshalBinary goes through them)."""
    if a is None:
        return True
    acc = (F.one, F.one, F.zero)
    base = to_jac(F, a)
    for bit in bin(R)[2:]:
        acc = jac_dbl(F, acc)
        if bit == "1":
            acc = jac_add(F, acc, base)
    return to_affine(F, acc) is None


def pt_sum(F, pts):
    acc = (F.one, F.one, F.zero)
    for p in pts:
        acc = jac_add(F, acc, to_jac(F, p))
    return to_affine(F, acc)


def g1_mul(k, a=G1_GEN): return pt_mul(F1, k, a)
def g2_mul(k, a=G2_GEN): return pt_mul(F2, k, a)
def g1_add(a, b): return pt_add(F1, a, b)
def g2_add(a, b): return pt_add(F2, a, b)


def msm_naive(F, scalars: Sequence[int], points) -> object:
    """Poly.BlindEval, algebra.go:348-359: sum_i p[i] * P[i], one scalar-mul per term."""
    if len(scalars) != len(points):
        raise ValueError("mismatch of length between poly %d and blinded eval points %d" % (len(scalars), len(points)))
    acc = (F.one, F.one, F.zero)
    for k, p in zip(scalars, points):
        q = pt_mul(F, k, p)
        acc = jac_add(F, acc, to_jac(F, q))
    return to_affine(F, acc)


def msm_fast(F, scalars: Sequence[int], points, c: int = 0):
    """Same group element as msm_naive, by a plain bucket method (checker speed-up only)."""
    n = len(scalars)
    assert n == len(points)
    if n == 0:
        return None
    if c == 0:
        c = max(2, min(16, n.bit_length() - 2))
    inf = (F.one, F.one, F.zero)
    nwin = (255 + c - 1) // c
    total = inf
    for w in reversed(range(nwin)):
        for _ in range(c):
            total = jac_dbl(F, total)
        buckets = {}
        for k, p in zip(scalars, points):
            d = ((k % R) >> (w * c)) & ((1 << c) - 1)
            if d and p is not None:
                buckets[d] = jac_add(F, buckets[d], to_jac(F, p)) if d in buckets else to_jac(F, p)
        run = inf; acc = inf
        for d in range((1 << c) - 1, 0, -1):
            if d in buckets:
                run = jac_add(F, run, buckets[d])
            acc = jac_add(F, acc, run) if (run[2] != F.zero) else acc
        total = jac_add(F, total, acc)
    return to_affine(F, total)


# --------------------------------------------------------------------------------------------
# Wire formats (kyber MarshalBinary; pinochio.go:256-275): Fr 32 B big-endian; G1 48 B / G2 96 B
# zcash-compressed (0x80 compressed, 0x40 infinity, 0x20 y is the lexicographically larger root;
# Fp2 is ordered by c1 first and serialised c1 || c0).
# --------------------------------------------------------------------------------------------
def fr_to_bytes(x: int) -> bytes:
    return (x % R).to_bytes(32, "big")


def fr_from_bytes(b: bytes) -> int:
    return int.from_bytes(b, "big")


def _fp_larger(y: int) -> bool:
    return y > (P - 1) // 2


def _fp2_larger(y) -> bool:
    return _fp_larger(y[1]) if y[1] != 0 else _fp_larger(y[0])


def g1_compress(pt) -> bytes:
    if pt is None:
        return bytes([0xC0]) + bytes(47)
    b = bytearray(pt[0].to_bytes(48, "big"))
    b[0] |= 0x80 | (0x20 if _fp_larger(pt[1]) else 0)
    return bytes(b)


def fp_sqrt(a: int) -> Optional[int]:
    s = pow(a, (P + 1) // 4, P)
    return s if s * s % P == a % P else None


def g1_decompress(b: bytes):
    assert len(b) == 48 and b[0] & 0x80
    if b[0] & 0x40:
        return None
    x = int.from_bytes(bytes([b[0] & 0x1F]) + b[1:], "big")
    y = fp_sqrt((x * x * x + 4) % P)
    if y is None:
        raise ValueError("bad G1 encoding")
    if _fp_larger(y) != bool(b[0] & 0x20):
        y = P - y
    if not in_subgroup(F1, (x, y)):
        raise ValueError("G1 point outside the prime-order subgroup")
    return (x, y)


def fp2_sqrt(a):
    """Square root in Fp2 (p = 3 mod 4), complex method; returns None for non-residues."""
    if a == (0, 0):
        return (0, 0)
    a1 = fp2_pow(a, (P - 3) // 4)
    alpha = F2.mul(F2.sqr(a1), a)
    x0 = F2.mul(a1, a)
    if alpha == (P - 1, 0):
        s = (F1.neg(x0[1]), x0[0])  # u * x0
    else:
        bb = fp2_pow(F2.add(alpha, F2.one), (P - 1) // 2)
        s = F2.mul(bb, x0)
    return s if F2.sqr(s) == a else None


def fp2_pow(a, e: int):
    r = F2.one
    for bit in bin(e)[2:]:
        r = F2.sqr(r)
        if bit == "1":
            r = F2.mul(r, a)
    return r


def g2_compress(pt) -> bytes:
    if pt is None:
        return bytes([0xC0]) + bytes(95)
    (x0, x1), y = pt
    b = bytearray(x1.to_bytes(48, "big") + x0.to_bytes(48, "big"))
    b[0] |= 0x80 | (0x20 if _fp2_larger(y) else 0)
    return bytes(b)


def g2_decompress(b: bytes):
    assert len(b) == 96 and b[0] & 0x80
    if b[0] & 0x40:
        return None
    x1 = int.from_bytes(bytes([b[0] & 0x1F]) + b[1:48], "big")
    x0 = int.from_bytes(b[48:], "big")
    x = (x0, x1)
    y = fp2_sqrt(F2.add(F2.mul(F2.sqr(x), x), F2.b))
    if y is None:
        raise ValueError("bad G2 encoding")
    if _fp2_larger(y) != bool(b[0] & 0x20):
        y = F2.neg(y)
    if not in_subgroup(F2, (x, y)):
        raise ValueError("G2 point outside the prime-order subgroup")
    return (x, y)


def g1_affine_bytes(pt) -> bytes:
    """zcash uncompressed: x || y big-endian, 96 B; infinity = 0x40 flag."""
    if pt is None:
        return bytes([0x40]) + bytes(95)
    return pt[0].to_bytes(48, "big") + pt[1].to_bytes(48, "big")


def g2_affine_bytes(pt) -> bytes:
    """zcash uncompressed: x.c1 || x.c0 || y.c1 || y.c0, 192 B."""
    if pt is None:
        return bytes([0x40]) + bytes(191)
    (x0, x1), (y0, y1) = pt
    return b"".join(v.to_bytes(48, "big") for v in (x1, x0, y1, y0))


def g1_from_affine_bytes(b: bytes):
    if b[0] & 0x40:
        return None
    return (int.from_bytes(b[:48], "big"), int.from_bytes(b[48:96], "big"))


def g2_from_affine_bytes(b: bytes):
    if b[0] & 0x40:
        return None
    x1, x0, y1, y0 = (int.from_bytes(b[i * 48:(i + 1) * 48], "big") for i in range(4))
    return ((x0, x1), (y0, y1))


# --------------------------------------------------------------------------------------------
# Deterministic sampling (replaces Pick(random.New()) so that runs are reproducible)
# --------------------------------------------------------------------------------------------
class Sampler:
    """SHA-256 counter stream reduced mod r, rejecting 0."""

    def __init__(self, seed: int = 0, tag: bytes = b"playsnark-b200"):
        self.seed, self.tag, self.ctr = seed, tag, 0

    def fr(self) -> int:
        while True:
            h = b"".join(
                hashlib.sha256(self.tag + self.seed.to_bytes(8, "big") + self.ctr.to_bytes(8, "big") + bytes([i])).digest()
                for i in range(2))
            self.ctr += 1
            v = int.from_bytes(h, "big") % R
            if v:
                return v


# --------------------------------------------------------------------------------------------
# Polynomials over Fr, coefficients low degree first (algebra.go:89-243)
# --------------------------------------------------------------------------------------------
Poly = List[int]


def value_to_fr(v: int) -> int:
    """Value.ToFieldElement, curve.go:17-19 (SetInt64 reduces mod r; negatives wrap)."""
    return v % R


def poly_mul(p: Poly, p2: Poly) -> Poly:
    """Poly.Mul, algebra.go:92-105 (schoolbook, length la+lb-1)."""
    out = [0] * (len(p) + len(p2) - 1)
    for i, v1 in enumerate(p):
        if v1 == 0:
            continue  # identical result; skips multiplications by zero
        for j, v2 in enumerate(p2):
            out[i + j] = (out[i + j] + v1 * v2) % R
    return out


def poly_add(p: Poly, p2: Poly) -> Poly:
    """Poly.Add, algebra.go:161-178."""
    out = [0] * max(len(p), len(p2))
    for i, v in enumerate(p):
        out[i] = v % R
    for i, v in enumerate(p2):
        out[i] = (out[i] + v) % R
    return out


def poly_sub(p: Poly, p2: Poly) -> Poly:
    """Poly.Sub, algebra.go:180-197."""
    out = [0] * max(len(p), len(p2))
    for i, v in enumerate(p):
        out[i] = v % R
    for i, v in enumerate(p2):
        out[i] = (out[i] - v) % R
    return out


def poly_eval(p: Poly, x: int) -> int:
    """Poly.Eval, algebra.go:107-115 (Horner from the top)."""
    v = 0
    for c in reversed(p):
        v = (v * x + c) % R
    return v


def poly_normalize(p: Poly) -> Poly:
    """Poly.Normalize, algebra.go:230-239."""
    n = len(p)
    while n > 0 and p[n - 1] % R == 0:
        n -= 1
    return p[:n]


def poly_div2(p: Poly, p2: Poly) -> Tuple[Poly, Poly]:
    """Poly.Div2, algebra.go:140-159: long division from the top.  The quotient has
    len(p)-len(p2)+1 entries whatever the leading zeros (first q.Add(tPoly) fixes its length);
    the remainder keeps len(p2)-1 entries (the in-function Normalize() result is discarded)."""
    r = [c % R for c in p]
    q: Poly = []
    lead_inv = pow(p2[-1], -1, R)
    while len(r) > 0 and len(r) >= len(p2):
        t = r[-1] * lead_inv % R
        deg_t = len(r) - len(p2)
        if len(q) < deg_t + 1:
            q = q + [0] * (deg_t + 1 - len(q))
        q[deg_t] = (q[deg_t] + t) % R
        # r = (r - tPoly*p2)[:len(r)-1]; tPoly is the monomial t*x^deg_t
        for j, c in enumerate(p2):
            r[deg_t + j] = (r[deg_t + j] - t * c) % R
        r = r[:-1]
    return q, r


def poly_div_synthetic(p: Poly, p2: Poly) -> Tuple[Poly, Poly]:
    """Poly.Div, algebra.go:119-137.  NOTE the reference routine treats slices as HIGH degree
    first (it divides out[i] by divisor[0]); restated as written, used only by its own KAT
    (algebra_test.go:48-74)."""
    out = [c % R for c in p]
    for i in range(len(p) - (len(p2) - 1)):
        out[i] = out[i] * pow(p2[0], -1, R) % R
        coef = out[i]
        if coef != 0:
            for j in range(1, len(p2)):
                out[i + j] = (out[i + j] + (-p2[j]) * coef) % R
    sep = len(out) - (len(p2) - 1)
    return out[:sep], out[sep:]


def lagrange_basis(i: int, xs: Sequence[int]) -> Poly:
    """lagrangeBasis, algebra.go:317-338 for x-coordinates xs (ints), index i into xs."""
    basis = [1]
    acc = 1
    for m, xm in enumerate(xs):
        if m == i:
            continue
        basis = poly_mul(basis, [(-xm) % R, 1])
        acc = acc * pow((xs[i] - xm) % R, -1, R) % R
    return [c * acc % R for c in basis]


def interpolate(ys: Sequence[int]) -> Poly:
    """Interpolate, algebra.go:254-280: p(1)=y_1 ... p(n)=y_n, faithful O(n^3) form."""
    xs = list(range(1, len(ys) + 1))
    acc = [0]
    for j in range(len(xs)):
        basis = [c * ys[j] % R for c in lagrange_basis(j, xs)]
        acc = poly_add(acc, basis)
    return acc


def vanishing_poly(n: int) -> Poly:
    """z(x) = prod_{i=1..n} (x - i), qap.go:41-55."""
    z = [1]
    for i in range(1, n + 1):
        # multiply by (x - i)
        nz = [0] * (len(z) + 1)
        for k, c in enumerate(z):
            nz[k] = (nz[k] - i * c) % R
            nz[k + 1] = (nz[k + 1] + c) % R
        z = nz
    return z


def lagrange_basis_all_fast(n: int) -> List[Poly]:
    """All l_j(x) on the domain {1..n} via l_j = z/((x-j) z'(j)); O(n^2) total.  Same polynomials
    as lagrange_basis(j-1, [1..n]) (uniqueness of the interpolant); checker speed-up only."""
    z = vanishing_poly(n)
    fact = [1] * (n + 1)
    for i in range(1, n + 1):
        fact[i] = fact[i - 1] * i % R
    out = []
    for j in range(1, n + 1):
        # synthetic division of z by (x - j)
        q = [0] * n
        carry = 0
        for k in range(n, 0, -1):
            carry = (z[k] + carry * j) % R
            q[k - 1] = carry
        zp = fact[j - 1] * fact[n - j] % R
        if (n - j) & 1:
            zp = (-zp) % R
        s = pow(zp, -1, R)
        out.append([c * s % R for c in q])
    return out


# --------------------------------------------------------------------------------------------
# Integer matrices / R1CS (algebra.go:11-87, r1cs.go)
# --------------------------------------------------------------------------------------------
def mat_transpose(m): return [list(col) for col in zip(*m)]   # Matrix.Transpose, algebra.go:42-49
def mat_mul_vec(m, v): return [sum(a * b for a, b in zip(row, v)) for row in m]  # Matrix.Mul :51-61
def hadamard(a, b): return [x * y for x, y in zip(a, b)]      # Vector.Hadamard :63-69


class R1CS:
    """r1cs.go:81-174.  Variable order [const, inputs..., outputs..., intermediates...] (:132-144)."""

    def __init__(self):
        self.inputs: List[str] = []
        self.outputs: List[str] = []
        self.intermediates: List[str] = []
        self.vars: List[str] = ["const"]
        self.left: List[List[int]] = []
        self.right: List[List[int]] = []
        self.out: List[List[int]] = []

    def nb_io(self) -> int:  # r1cs.go:108-110
        return 1 + len(self.inputs) + len(self.outputs)

    def _merge(self):  # r1cs.go:132-144
        self.vars = ["const"] + self.inputs + self.outputs + self.intermediates

    def new_input(self, name): self.inputs.append(name); self._merge()
    def new_output(self, name): self.outputs.append(name); self._merge()
    def new_var(self, name): self.intermediates.append(name); self._merge()

    def index_of(self, name):  # r1cs.go:21-28
        if name not in self.vars:
            raise KeyError("plouf")
        return self.vars.index(name)

    def constraint_on(self, *names):  # r1cs.go:33-50
        return [1 if v in names else 0 for v in self.vars]

    def mul(self, left, right, out):  # r1cs.go:148-152
        self.left.append(self.constraint_on(left)); self.right.append(self.constraint_on(right))
        self.out.append(self.constraint_on(out))

    def add(self, v1, v2, out):  # r1cs.go:156-164
        self.left.append(self.constraint_on(v1, v2)); self.right.append(self.constraint_on("const"))
        self.out.append(self.constraint_on(out))

    def add_const(self, v1, add, out):  # r1cs.go:168-174
        row = self.constraint_on("const", v1)
        row[0] = row[0] * add
        self.left.append(row); self.right.append(self.constraint_on("const"))
        self.out.append(self.constraint_on(out))


def create_r1cs() -> R1CS:
    """createR1CS, r1cs.go:178-198: x^3 + x + 5 = 35."""
    c = R1CS()
    c.new_input("x"); c.new_output("out")
    c.new_var("u"); c.new_var("v"); c.new_var("w")
    c.mul("x", "x", "u"); c.mul("u", "x", "v"); c.add("v", "x", "w"); c.add_const("w", 5, "out")
    return c


def create_witness(r: R1CS) -> List[int]:
    """createWitness, r1cs.go:67-76."""
    sol = [0] * len(r.vars)
    for name, val in (("const", 1), ("x", 3), ("out", 35), ("u", 9), ("v", 27), ("w", 30)):
        sol[r.index_of(name)] = val
    return sol


# --------------------------------------------------------------------------------------------
# QAP (qap.go)
# --------------------------------------------------------------------------------------------
class QAP:
    """qap.go:10-27."""

    def __init__(self, nb_vars, nb_io, nb_gates, left, right, out, z):
        self.nb_vars, self.nb_io, self.nb_gates = nb_vars, nb_io, nb_gates
        self.left, self.right, self.out, self.z = left, right, out, z

    def compute_aggregate_poly(self, sol: Sequence[int]):
        """computeAggregatePoly, qap.go:164-175; `sol` entries are Fr values (ints mod r) or Go ints."""
        left: Poly = []; right: Poly = []; out: Poly = []
        for i, val in enumerate(sol):
            pv = [value_to_fr(val)]
            left = poly_add(left, poly_mul(self.left[i], pv))
            right = poly_add(right, poly_mul(self.right[i], pv))
            out = poly_add(out, poly_mul(self.out[i], pv))
        return left, right, out

    def sanity_check(self, sol):  # qap.go:177-189
        if len(sol) != len(self.left):
            raise ValueError("different number of solution variables than left polynomials")
        if len(sol) != len(self.right):
            raise ValueError("different numberof solution variables than right polynomials")
        if len(sol) != len(self.out):
            raise ValueError("different numbers of solutions variables than out polynomials")

    def is_valid(self, sol) -> bool:  # qap.go:107-149
        self.sanity_check(sol)
        l, r, o = self.compute_aggregate_poly(sol)
        _, rem = poly_div2(poly_sub(poly_mul(l, r), o), self.z)
        return len(poly_normalize(rem)) == 0

    def quotient(self, sol) -> Poly:  # qap.go:151-162
        l, r, o = self.compute_aggregate_poly(sol)
        h, rem = poly_div2(poly_sub(poly_mul(l, r), o), self.z)
        if len(poly_normalize(rem)) > 0:
            raise ArithmeticError("apocalypse")
        return h


def to_qap(circuit: R1CS, fast: bool = True) -> QAP:
    """ToQAP, qap.go:35-65 + qapInterpolate :67-93.  fast=False follows Interpolate literally
    (O(n^3) per variable); fast=True builds the same polynomials from the Lagrange basis."""
    n = len(circuit.left)

    def interp_all(m):
        cols = mat_transpose(m)
        if not fast:
            return [interpolate([value_to_fr(v) for v in col]) for col in cols]
        basis = lagrange_basis_all_fast(n)
        res = []
        for col in cols:
            acc = [0] * n
            for j, v in enumerate(col):
                if v:
                    fv = value_to_fr(v)
                    bj = basis[j]
                    for k in range(n):
                        acc[k] = (acc[k] + fv * bj[k]) % R
            res.append(acc)
        return res

    return QAP(len(circuit.vars), circuit.nb_io(), n, interp_all(circuit.left), interp_all(circuit.right),
               interp_all(circuit.out), vanishing_poly(n))


def blind_eval(F, p: Poly, pts):
    """Poly.BlindEval, algebra.go:348-359 (exact length match or panic)."""
    return msm_naive(F, p, pts)


def generate_powers_commit(F, gen, e: int, shift: int, power: int):
    """GeneratePowersCommit, algebra.go:371-384: { (shift * e^i) * generator }, i = 0..power."""
    out = []
    si = 1
    out.append(pt_mul(F, shift, gen))
    for _ in range(power):
        si = si * e % R
        out.append(pt_mul(F, si * shift % R, gen))
    return out


# --------------------------------------------------------------------------------------------
# Groth16 (groth16.go)
# --------------------------------------------------------------------------------------------
class Groth16Setup:
    pass


def linear_poly_for_var(qap: QAP, i, x, alpha, beta) -> int:
    """linearPolyForVar, groth16.go:238-249."""
    ui = poly_eval(qap.left[i], x); vi = poly_eval(qap.right[i], x); wi = poly_eval(qap.out[i], x)
    return (wi + beta * ui + alpha * vi) % R


def groth16_setup(qap: QAP, sampler: Sampler, mul_g1=g1_mul, mul_g2=g2_mul) -> Groth16Setup:
    """NewGroth16TrustedSetup, groth16.go:64-101, sampling order preserved
    (alpha, beta, delta, x, gamma)."""
    tr = Groth16Setup()
    tw = {}
    tw["Alpha"] = sampler.fr(); tr.Alpha = mul_g1(tw["Alpha"])
    tw["Beta"] = sampler.fr(); tr.Beta = mul_g1(tw["Beta"]); tr.Beta2 = mul_g2(tw["Beta"])
    tw["Delta"] = sampler.fr(); tr.Delta = mul_g1(tw["Delta"]); tr.Delta2 = mul_g2(tw["Delta"])
    tw["X"] = sampler.fr()
    x = tw["X"]
    n = qap.nb_gates
    pw = [pow(x, i, R) for i in range(n)]
    tr.Xi = [mul_g1(e) for e in pw]
    tr.Xi2 = [mul_g2(e) for e in pw]
    tw["Gamma"] = sampler.fr(); tr.Gamma = mul_g2(tw["Gamma"])
    diff = qap.nb_vars - qap.nb_io
    def full(lo, hi, div):
        dinv = pow(div, -1, R)
        lps = [linear_poly_for_var(qap, i, x, tw["Alpha"], tw["Beta"]) * dinv % R for i in range(lo, hi)]
        return lps, [mul_g1(lp) for lp in lps]
    tw["IoLP"], tr.IoLP = full(0, diff, tw["Gamma"])
    tw["NioLP"], tr.NioLP = full(diff, qap.nb_vars, tw["Delta"])
    txd = poly_eval(qap.z, x) * pow(tw["Delta"], -1, R) % R
    tr.XiT = [mul_g1(pw[i] * txd % R) for i in range(n - 1)]
    tr.tw = tw
    return tr


def groth16_prove(tr: Groth16Setup, q: QAP, sol: Sequence[int], r: int, s: int, faithful: bool = False):
    """Groth16Prove, groth16.go:122-211 with the blinding scalars injected.  faithful=True walks
    sumBlind variable by variable (m*n scalar-muls, :134-141); otherwise the closed form
    A = Alpha + sum_k a_k Xi[k] + r Delta etc. (same group elements)."""
    diff = q.nb_vars - q.nb_io
    fe = [value_to_fr(v) for v in sol]

    def sum_blind(F, polys, xi):
        if faithful:
            acc = None
            for i in range(q.nb_vars):
                uix = blind_eval(F, polys[i], xi)
                acc = pt_add(F, acc, pt_mul(F, fe[i], uix))
            return acc
        agg: Poly = []
        for i in range(q.nb_vars):
            agg = poly_add(agg, poly_mul(polys[i], [fe[i]]))
        return msm_fast(F, agg, xi)

    A = sum_blind(F1, q.left, tr.Xi)
    A = g1_add(A, g1_mul(r, tr.Delta))
    A = g1_add(tr.Alpha, A)
    B = sum_blind(F2, q.right, tr.Xi2)
    B = g2_add(B, g2_mul(s, tr.Delta2))
    B = g2_add(tr.Beta2, B)
    nio = None
    for i, pt in enumerate(tr.NioLP):
        nio = g1_add(nio, g1_mul(fe[i + diff], pt))
    C = nio
    h = q.quotient(sol)
    htd = blind_eval(F1, h, tr.XiT) if faithful else (
        msm_fast(F1, h, tr.XiT) if len(h) == len(tr.XiT) else blind_eval(F1, h, tr.XiT))
    C = g1_add(C, htd)
    C = g1_add(C, g1_mul(s, A))
    B1 = sum_blind(F1, q.right, tr.Xi)
    B1 = g1_add(B1, g1_mul(s, tr.Delta))
    B1 = g1_add(B1, tr.Beta)
    C = g1_add(C, g1_mul(r, B1))
    rsd = g1_mul(r * s % R, tr.Delta)
    C = g1_add(C, pt_neg(F1, rsd))
    return {"A": A, "B": B, "C": C, "R": r, "S": s, "h": h}


def groth16_expected_from_toxic(tr: Groth16Setup, q: QAP, sol, r: int, s: int):
    """TestGroth16ProofGen, groth16_test.go:32-107: recompute A, B, C in the exponent from the
    toxic waste, one scalar-mul per element."""
    tw = tr.tw
    x = tw["X"]
    fe = [value_to_fr(v) for v in sol]
    diff = q.nb_vars - q.nb_io
    ea = sum(poly_eval(q.left[i], x) * fe[i] for i in range(q.nb_vars)) % R
    ea = (ea + r * tw["Delta"] + tw["Alpha"]) % R
    eb = sum(poly_eval(q.right[i], x) * fe[i] for i in range(q.nb_vars)) % R
    eb = (eb + s * tw["Delta"] + tw["Beta"]) % R
    dinv = pow(tw["Delta"], -1, R)
    ec = 0
    for i in range(diff, q.nb_vars):
        ec += linear_poly_for_var(q, i, x, tw["Alpha"], tw["Beta"]) * fe[i] % R * dinv
    h = q.quotient(sol)
    ec += poly_eval(h, x) * poly_eval(q.z, x) % R * dinv
    ec += s * ea + r * eb - r * s % R * tw["Delta"]
    return g1_mul(ea % R), g2_mul(eb % R), g1_mul(ec % R)


def groth16_verify(tr: Groth16Setup, q: QAP, proof, io: Sequence[int]) -> bool:
    """Groth16Verify, groth16.go:214-233 (GT 'Add' is Fp12 multiplication)."""
    left = pairing(proof["A"], proof["B"])
    a = pairing(tr.Alpha, tr.Beta2)
    b1 = None
    for i, iolp in enumerate(tr.IoLP):
        b1 = g1_add(b1, g1_mul(value_to_fr(io[i]), iolp))
    b = pairing(b1, tr.Gamma)
    c = pairing(proof["C"], tr.Delta2)
    return left == fp12_mul(a, fp12_mul(b, c))


# --------------------------------------------------------------------------------------------
# PHGR13 / Pinocchio (pinochio.go)
# --------------------------------------------------------------------------------------------
def generate_eval_commit(F, base, polys, x, shift):
    """generateEvalCommit, pinochio.go:381-388."""
    return [pt_mul(F, poly_eval(poly_normalize(p), x) * shift % R, base) for p in polys]


def phgr13_setup(qap: QAP, sampler: Sampler, with_vk: bool = True):
    """NewPHGR13TrustedSetup, pinochio.go:93-176, sampling order preserved
    (s, av, aw, ay, rv, rw, beta, gamma).  with_vk=False skips the three all-variable commitment
    vectors of the verification key (only needed by PHGR13Verify) for large prover-only tests."""
    ek, vk, t = {}, {}, {}
    s = sampler.fr()
    ek["gsi"] = generate_powers_commit(F1, G1_GEN, s, 1, (len(qap.z) - 1) - 2)
    av, aw, ay = sampler.fr(), sampler.fr(), sampler.fr()
    rv = sampler.fr(); gv = g1_mul(rv)
    rw = sampler.fr(); gw = g2_mul(rw); g1w = g1_mul(rw)
    ry = rv * rw % R; gy = g1_mul(ry); g2y = g2_mul(ry)
    diff = qap.nb_vars - qap.nb_io
    ek["vs"] = generate_eval_commit(F1, gv, qap.left[diff:], s, 1)
    ek["ws"] = generate_eval_commit(F2, gw, qap.right[diff:], s, 1)
    ek["ys"] = generate_eval_commit(F1, gy, qap.out[diff:], s, 1)
    ek["vas"] = generate_eval_commit(F1, gv, qap.left[diff:], s, av)
    ek["was"] = generate_eval_commit(F1, g1w, qap.right[diff:], s, aw)
    ek["yas"] = generate_eval_commit(F1, gy, qap.out[diff:], s, ay)
    beta = sampler.fr()
    ek["vbs"] = generate_eval_commit(F1, gv, qap.left[diff:], s, beta)
    ek["wbs"] = generate_eval_commit(F1, g1w, qap.right[diff:], s, beta)
    ek["ybs"] = generate_eval_commit(F1, gy, qap.out[diff:], s, beta)
    gamma = sampler.fr()
    bgamma = gamma * beta % R
    vk["g1"] = G1_GEN
    vk["av"] = g2_mul(av); vk["aw"] = g1_mul(aw); vk["ay"] = g2_mul(ay)
    vk["gamma"] = g2_mul(gamma); vk["bgamma"] = g1_mul(bgamma); vk["bgamma2"] = g2_mul(bgamma)
    vk["yts"] = g2_mul(poly_eval(qap.z, s), g2y)
    if with_vk:
        vk["vs"] = generate_eval_commit(F1, gv, qap.left, s, 1)
        vk["ws"] = generate_eval_commit(F2, gw, qap.right, s, 1)
        vk["ys"] = generate_eval_commit(F1, gy, qap.out, s, 1)
    t.update(beta=beta, s=s, gv=gv, gw=gw, gy=gy, ry=ry, rv=rv, rw=rw)
    return {"EK": ek, "VK": vk, "t": t}


PHGR13_FIELDS = ("hs", "vss", "wss", "yss", "vass", "wass", "yass", "gz")


def phgr13_prove(ek, qap: QAP, solution):
    """PHGR13Prove, pinochio.go:207-254."""
    l, r, o = qap.compute_aggregate_poly(solution)
    hx, rem = poly_div2(poly_sub(poly_mul(l, r), o), qap.z)
    if len(poly_normalize(rem)) > 0:
        raise ArithmeticError("apocalypse")
    ghs = blind_eval(F1, hx, ek["gsi"])
    diff = qap.nb_vars - qap.nb_io
    fe = [value_to_fr(v) for v in solution]

    def sol_commit(F, ec):  # computeSolCommit, pinochio.go:222-229
        acc = None
        for i, e in enumerate(ec):
            acc = pt_add(F, acc, pt_mul(F, fe[diff + i], e))
        return acc

    gvb, gwb, gyb = sol_commit(F1, ek["vbs"]), sol_commit(F1, ek["wbs"]), sol_commit(F1, ek["ybs"])
    return {
        "hs": ghs, "vss": sol_commit(F1, ek["vs"]), "wss": sol_commit(F2, ek["ws"]),
        "yss": sol_commit(F1, ek["ys"]), "vass": sol_commit(F1, ek["vas"]),
        "wass": sol_commit(F1, ek["was"]), "yass": sol_commit(F1, ek["yas"]),
        "gz": g1_add(gvb, g1_add(gwb, gyb)), "h": hx,
    }


def compute_commit_io_solution(F, poly, io):
    """computeCommitIOSolution, pinochio.go:390-407."""
    acc = None
    for i, gs in enumerate(poly):
        acc = pt_add(F, acc, pt_mul(F, value_to_fr(io[i]), gs))
    return acc


def phgr13_verify(vk, qap: QAP, p, io) -> bool:
    """PHGR13Verify, pinochio.go:281-375."""
    diff = qap.nb_vars - qap.nb_io
    gv = g1_add(compute_commit_io_solution(F1, vk["vs"][:diff], io), p["vss"])
    gw = g2_add(compute_commit_io_solution(F2, vk["ws"][:diff], io), p["wss"])
    gy = g1_add(compute_commit_io_solution(F1, vk["ys"][:diff], io), p["yss"])
    left = pairing(gv, gw)
    right = fp12_mul(pairing(p["hs"], vk["yts"]), pairing(gy, G2_GEN))
    if left != right:
        return False
    if pairing(p["vass"], G2_GEN) != pairing(p["vss"], vk["av"]):
        return False
    if pairing(p["wass"], G2_GEN) != pairing(vk["aw"], p["wss"]):
        return False
    if pairing(p["yass"], G2_GEN) != pairing(p["yss"], vk["ay"]):
        return False
    left = pairing(p["gz"], vk["gamma"])
    t1 = pairing(g1_add(p["vss"], p["yss"]), vk["bgamma2"])
    t2 = pairing(vk["bgamma"], p["wss"])
    return fp12_mul(t1, t2) == left


# --------------------------------------------------------------------------------------------
# Pairing: optimal ate on BLS12-381, tower Fp2 -> Fp6 = Fp2[v]/(v^3 - xi) -> Fp12 = Fp6[w]/(w^2 - v),
# xi = 1 + u.  Suite.Pair, curve.go:36-38.  Only equality of GT values is ever observed by the
# reference (groth16.go:232, pinochio.go:319-372), which any non-degenerate bilinear map decides
# identically.
# --------------------------------------------------------------------------------------------
def _mul_xi(a):  # (a0 + a1 u)(1 + u)
    return ((a[0] - a[1]) % P, (a[0] + a[1]) % P)


FP6_ZERO = (F2.zero, F2.zero, F2.zero)
FP6_ONE = (F2.one, F2.zero, F2.zero)
FP12_ONE = (FP6_ONE, FP6_ZERO)


def fp6_add(a, b): return tuple(F2.add(x, y) for x, y in zip(a, b))
def fp6_sub(a, b): return tuple(F2.sub(x, y) for x, y in zip(a, b))
def fp6_neg(a): return tuple(F2.neg(x) for x in a)


def fp6_mul(a, b):
    a0, a1, a2 = a; b0, b1, b2 = b
    t0, t1, t2 = F2.mul(a0, b0), F2.mul(a1, b1), F2.mul(a2, b2)
    c0 = F2.add(t0, _mul_xi(F2.sub(F2.mul(F2.add(a1, a2), F2.add(b1, b2)), F2.add(t1, t2))))
    c1 = F2.add(F2.sub(F2.mul(F2.add(a0, a1), F2.add(b0, b1)), F2.add(t0, t1)), _mul_xi(t2))
    c2 = F2.add(F2.sub(F2.mul(F2.add(a0, a2), F2.add(b0, b2)), F2.add(t0, t2)), t1)
    return (c0, c1, c2)


def fp6_mul_v(a):  # multiply by v
    return (_mul_xi(a[2]), a[0], a[1])


def fp6_inv(a):
    a0, a1, a2 = a
    c0 = F2.sub(F2.sqr(a0), _mul_xi(F2.mul(a1, a2)))
    c1 = F2.sub(_mul_xi(F2.sqr(a2)), F2.mul(a0, a1))
    c2 = F2.sub(F2.sqr(a1), F2.mul(a0, a2))
    t = F2.add(F2.mul(a0, c0), _mul_xi(F2.add(F2.mul(a2, c1), F2.mul(a1, c2))))
    ti = F2.inv(t)
    return (F2.mul(c0, ti), F2.mul(c1, ti), F2.mul(c2, ti))


def fp12_mul(a, b):
    a0, a1 = a; b0, b1 = b
    t0, t1 = fp6_mul(a0, b0), fp6_mul(a1, b1)
    c1 = fp6_sub(fp6_mul(fp6_add(a0, a1), fp6_add(b0, b1)), fp6_add(t0, t1))
    return (fp6_add(t0, fp6_mul_v(t1)), c1)


def fp12_sqr(a): return fp12_mul(a, a)
def fp12_conj(a): return (a[0], fp6_neg(a[1]))


def fp12_inv(a):
    a0, a1 = a
    t = fp6_inv(fp6_sub(fp6_mul(a0, a0), fp6_mul_v(fp6_mul(a1, a1))))
    return (fp6_mul(a0, t), fp6_neg(fp6_mul(a1, t)))


def fp12_pow(a, e: int):
    r = FP12_ONE
    for bit in bin(e)[2:]:
        r = fp12_sqr(r)
        if bit == "1":
            r = fp12_mul(r, a)
    return r


def _line(lam, xt, yt, xp, yp):
    """Line through the (untwisted) G2 point with twisted slope lam, evaluated at P=(xp,yp) and
    scaled by w^3 (a factor the final exponentiation removes):
    (lam*xt - yt) + (-lam*xp) v + (yp) v w."""
    A = F2.sub(F2.mul(lam, xt), yt)
    B = F2.muli(F2.neg(lam), xp)
    return ((A, B, F2.zero), (F2.zero, (yp % P, 0), F2.zero))


def miller_loop(Pt, Q):
    if Pt is None or Q is None:
        return FP12_ONE
    xp, yp = Pt
    T = Q
    f = FP12_ONE
    for bit in bin(BLS_X)[3:]:
        xt, yt = T
        lam = F2.mul(F2.muli(F2.sqr(xt), 3), F2.inv(F2.muli(yt, 2)))
        f = fp12_mul(fp12_sqr(f), _line(lam, xt, yt, xp, yp))
        x3 = F2.sub(F2.sqr(lam), F2.muli(xt, 2))
        T = (x3, F2.sub(F2.mul(lam, F2.sub(xt, x3)), yt))
        if bit == "1":
            xt, yt = T
            lam = F2.mul(F2.sub(Q[1], yt), F2.inv(F2.sub(Q[0], xt)))
            f = fp12_mul(f, _line(lam, xt, yt, xp, yp))
            x3 = F2.sub(F2.sub(F2.sqr(lam), xt), Q[0])
            T = (x3, F2.sub(F2.mul(lam, F2.sub(xt, x3)), yt))
    return fp12_conj(f)  # the curve parameter is negative


_HARD_EXP = (P ** 2 + 1) * ((P ** 4 - P ** 2 + 1) // R)


def final_exp(f):
    f1 = fp12_mul(fp12_conj(f), fp12_inv(f))  # f^(p^6 - 1)
    return fp12_pow(f1, _HARD_EXP)


def pairing(Pt, Q):
    """e(P, Q), P in G1 (affine over Fp), Q in G2 (affine over Fp2 on the twist)."""
    return final_exp(miller_loop(Pt, Q))
