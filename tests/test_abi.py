"""The product shared object loads and exports every symbol include/playsnark_b200.h declares; no
compute call is made (no GPU here)."""
import os
import re

from playsnark_b200 import _lib as L, build as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "playsnark_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ps_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_binding_agree():
    assert header_symbols() == sorted(name for name, _, _ in L.SYMBOLS)


def test_library_exports_every_symbol():
    B.build_cuda()
    lib = L.load()
    for name in header_symbols():
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.ps_version()
    assert lib.ps_strerror(L.PS_ERR_REMAINDER) == b"apocalypse"


def test_product_has_no_cpu_fallback():
    """nothing under playsnark_b200/ imports the oracle or the host emulation implicitly"""
    pkg = os.path.join(ROOT, "playsnark_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("# noqa", ""), fn
    assert "libps_hostemu" not in open(os.path.join(pkg, "_lib.py")).read()
