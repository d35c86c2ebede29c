"""ctypes wrapper of oracle/ps_oracle.c (TEST INFRASTRUCTURE; see that file's header)."""
import ctypes as C

from . import build_oracle
from . import ps_oracle as O

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_oracle.build())
        _lib.oc_blind_eval_g1.restype = C.c_long
        _lib.oc_blind_eval_g1.argtypes = [C.c_char_p, C.c_char_p, C.c_long, C.c_char_p]
        _lib.oc_quotient.restype = C.c_int
        _lib.oc_quotient.argtypes = [C.c_char_p] * 4 + [C.c_long, C.c_int, C.c_char_p]
        _lib.oc_aggregate.argtypes = [C.c_char_p, C.c_char_p, C.c_long, C.c_long, C.c_char_p]
        _lib.oc_fr_dot.restype = None
        _lib.oc_fr_dot.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_char_p]
    return _lib


def blind_eval_g1(points_affine_bytes: bytes, scalars_be: bytes):
    """Poly.BlindEval over G1; returns the affine point (oracle tuple) or None."""
    n = len(scalars_be) // 32
    assert len(points_affine_bytes) == 96 * n
    out = C.create_string_buffer(96)
    lib().oc_blind_eval_g1(points_affine_bytes, scalars_be, n, out)
    return O.g1_from_affine_bytes(out.raw)


def g1_mul(pt, k: int):
    out = C.create_string_buffer(96)
    lib().oc_g1_mul(O.g1_affine_bytes(pt), O.fr_to_bytes(k), out)
    return O.g1_from_affine_bytes(out.raw)


def fp_mul(a: int, b: int) -> int:
    out = C.create_string_buffer(48)
    lib().oc_fp_mul(a.to_bytes(48, "big"), b.to_bytes(48, "big"), out)
    return int.from_bytes(out.raw, "big")


def fr_mul(a: int, b: int) -> int:
    out = C.create_string_buffer(32)
    lib().oc_fr_mul(a.to_bytes(32, "big"), b.to_bytes(32, "big"), out)
    return int.from_bytes(out.raw, "big")


def fr_inv(a: int) -> int:
    out = C.create_string_buffer(32)
    lib().oc_fr_inv(a.to_bytes(32, "big"), out)
    return int.from_bytes(out.raw, "big")


def _poly_bytes(p):
    return b"".join(O.fr_to_bytes(v) for v in p)


def aggregate(polys, witness):
    m, n = len(polys), len(polys[0])
    out = C.create_string_buffer(32 * n)
    lib().oc_aggregate(b"".join(_poly_bytes(p) for p in polys), _poly_bytes(witness), m, n, out)
    return [int.from_bytes(out.raw[32 * i:32 * i + 32], "big") for i in range(n)]


def quotient(a, b, c, z, faithful=False):
    """(a*b - c) / z via Poly.Mul / Sub / Div2; raises ArithmeticError("apocalypse") on a remainder."""
    n = len(a)
    assert len(b) == n and len(c) == n and len(z) == n + 1
    out = C.create_string_buffer(32 * max(1, n - 1))
    rc = lib().oc_quotient(_poly_bytes(a), _poly_bytes(b), _poly_bytes(c), _poly_bytes(z), n, 1 if faithful else 0, out)
    if rc:
        raise ArithmeticError("apocalypse")
    return [int.from_bytes(out.raw[32 * i:32 * i + 32], "big") for i in range(n - 1)]


def fr_dot(a_be, b_be) -> int:
    """sum_i a[i]*b[i] mod r; a_be, b_be: bytes or C-contiguous numpy uint8 arrays of n x 32 big-endian bytes"""
    def ptr(x):
        if isinstance(x, (bytes, bytearray)):
            return C.cast(C.c_char_p(bytes(x)), C.c_void_p), len(x)
        return C.c_void_p(x.ctypes.data), x.nbytes
    (pa, la), (pb, lb) = ptr(a_be), ptr(b_be)
    assert la == lb and la % 32 == 0
    out = C.create_string_buffer(32)
    lib().oc_fr_dot(pa, pb, la // 32, out)
    return int.from_bytes(out.raw, "big")


# ---- the reference's whole proving flow restated in C (oracle/ps_prover.c) ---------------------------
_gen_set = False


def _prover_lib():
    global _gen_set
    L = lib()
    if not _gen_set:
        L.op_set_generators(O.g1_affine_bytes(O.G1_GEN), O.g2_affine_bytes(O.G2_GEN))
        L.op_groth16_flow.restype = C.c_int
        L.op_phgr13_flow.restype = C.c_int
        L.op_blind_eval_g1_mt.restype = C.c_long
        L.op_blind_eval_g1_mt.argtypes = [C.c_char_p, C.c_char_p, C.c_long, C.c_int, C.c_char_p]
        _gen_set = True
    return L


def max_threads() -> int:
    return int(lib().op_max_threads())


def _matrix(rows, n, m):
    flat = (C.c_long * (n * m))()
    for j, row in enumerate(rows):
        for i, v in enumerate(row):
            if v:
                flat[j * m + i] = v
    return flat


def groth16_flow(r1cs, witness_fr, toxic, r: int, s: int, threads: int = 1, fast_qap: bool = False):
    """ToQAP -> NewGroth16TrustedSetup -> Groth16Prove in C on `r1cs` (oracle R1CS with dense Go-int rows).
    toxic = (alpha, beta, delta, x).  Returns (A, B, C oracle points, h, seconds dict)."""
    L = _prover_lib()
    n, m = len(r1cs.left), len(r1cs.vars)
    out = C.create_string_buffer(384)
    hb = C.create_string_buffer(32 * max(1, n - 1))
    sec = (C.c_double * 3)()
    rc = L.op_groth16_flow(_matrix(r1cs.left, n, m), _matrix(r1cs.right, n, m), _matrix(r1cs.out, n, m), C.c_long(n), C.c_long(m),
                           C.c_long(r1cs.nb_io()), _poly_bytes(witness_fr), b"".join(O.fr_to_bytes(t) for t in toxic),
                           O.fr_to_bytes(r), O.fr_to_bytes(s), C.c_int(threads), C.c_int(1 if fast_qap else 0), out, hb, sec)
    if rc == 1:
        raise ArithmeticError("apocalypse")
    assert rc == 0
    raw = out.raw
    h = [int.from_bytes(hb.raw[32 * i:32 * i + 32], "big") for i in range(n - 1)]
    return (O.g1_from_affine_bytes(raw[:96]), O.g2_from_affine_bytes(raw[96:288]), O.g1_from_affine_bytes(raw[288:]), h,
            {"to_qap_s": sec[0], "setup_s": sec[1], "prove_s": sec[2]})


def phgr13_flow(r1cs, witness_fr, toxic, threads: int = 1, fast_qap: bool = False):
    """ToQAP -> NewPHGR13TrustedSetup (evaluation key) -> PHGR13Prove in C.
    toxic = (s, av, aw, ay, rv, rw, beta).  Returns (dict of the eight proof points, h, seconds dict)."""
    L = _prover_lib()
    n, m = len(r1cs.left), len(r1cs.vars)
    out = C.create_string_buffer(864)
    hb = C.create_string_buffer(32 * max(1, n - 1))
    sec = (C.c_double * 3)()
    rc = L.op_phgr13_flow(_matrix(r1cs.left, n, m), _matrix(r1cs.right, n, m), _matrix(r1cs.out, n, m), C.c_long(n), C.c_long(m),
                          C.c_long(r1cs.nb_io()), _poly_bytes(witness_fr), b"".join(O.fr_to_bytes(t) for t in toxic),
                          C.c_int(threads), C.c_int(1 if fast_qap else 0), out, hb, sec)
    if rc == 1:
        raise ArithmeticError("apocalypse")
    assert rc == 0
    raw = out.raw
    names = ("hs", "vss", "yss", "vass", "wass", "yass", "gz")
    proof = {nm: O.g1_from_affine_bytes(raw[96 * i:96 * i + 96]) for i, nm in enumerate(names)}
    proof["wss"] = O.g2_from_affine_bytes(raw[672:864])
    h = [int.from_bytes(hb.raw[32 * i:32 * i + 32], "big") for i in range(n - 1)]
    return proof, h, {"to_qap_s": sec[0], "setup_s": sec[1], "prove_s": sec[2]}


def blind_eval_g1_mt(points_affine_bytes: bytes, scalars_be: bytes, threads: int):
    n = len(scalars_be) // 32
    assert len(points_affine_bytes) == 96 * n
    out = C.create_string_buffer(96)
    _prover_lib().op_blind_eval_g1_mt(points_affine_bytes, scalars_be, n, threads, out)
    return O.g1_from_affine_bytes(out.raw)
