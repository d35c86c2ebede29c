"""C++ host mirror (playsnark_b200/host/playsnark.hpp) on the README circuit: linked against the
host-emulation build on CPU and, under -m gpu, against the product library."""
import os
import subprocess

import pytest

from oracle import ps_oracle as O
from playsnark_b200 import build as B
from tests.parity_cases import gold

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "_build")


def write_tokens(path):
    g = gold("readme_circuit")
    out = []
    out.append("%d %d %d" % (g["nb_vars"], g["nb_io"], g["nb_gates"]))
    for key in ("left", "right", "out"):
        for p in g[key]:
            out.append("%d %s" % (len(p), " ".join(p)))
    out.append("%d %s" % (len(g["z"]), " ".join(g["z"])))
    out.append("%d %s" % (len(g["witness"]), " ".join(str(v) for v in g["witness"])))
    k = g["groth16"]
    out.append(" ".join(k[n] for n in ("Alpha", "Beta", "Delta", "Beta2", "Delta2")))
    for n in ("Xi", "Xi2", "XiT", "NioLP"):
        out.append("%d %s" % (len(k[n]), " ".join(k[n])))
    out.append(k["r"] + " " + k["s"])
    ek = g["phgr13"]["ek"]
    for n in ("gsi", "vs", "ws", "ys", "vas", "was", "yas", "vbs", "wbs", "ybs"):
        out.append("%d %s" % (len(ek[n]), " ".join(ek[n])))
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")
    return g


def run_against(libpath, exe_name):
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, exe_name)
    src = os.path.join(ROOT, "tests", "host_cpp_test.cpp")
    libdir, libfile = os.path.split(libpath)
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", exe, src, "-L" + libdir, "-l:" + libfile, "-Wl,-rpath," + libdir])
    tok = os.path.join(BUILD, "readme_tokens.txt")
    g = write_tokens(tok)
    res = subprocess.run([exe, tok], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr
    lines = [l.split() for l in res.stdout.strip().split("\n")]
    got = {}
    for name, val in lines:
        got.setdefault(name, []).append(val)
    assert got["h"] == g["h"]
    assert got["A"] == [g["groth16"]["A"]] and got["B"] == [g["groth16"]["B"]] and got["C"] == [g["groth16"]["C"]]
    for f in O.PHGR13_FIELDS:
        assert got[f] == [g["phgr13"]["proof"][f]], f
    # BlindEval(h, XiT) = h(x) t(x)/delta * G, the htd term of groth16.go:185
    I = lambda xs: [int(x, 16) for x in xs]
    pts = [O.g1_decompress(bytes.fromhex(p)) for p in g["groth16"]["XiT"]]
    assert got["htd"] == [O.g1_compress(O.msm_naive(O.F1, I(g["h"]), pts)).hex()]
    assert got["err"] == ["apocalypse"] and got["len"] == ["mismatch"]
    assert got["neg"] == ["%064x" % (O.R - 1)]


def test_cpp_host_mirror_emulated():
    run_against(B.build_host_emulation(BUILD), "host_cpp_test_emu")


@pytest.mark.gpu
def test_cpp_host_mirror_gpu():
    run_against(B.LIB, "host_cpp_test_gpu")
