"""Builds libplaysnark_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libplaysnark_b200.so")

CU_SOURCES = [os.path.join(CSRC, "capi.cu"), os.path.join(CSRC, "accum_g2.cu")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "--expt-relaxed-constexpr",
    "-lineinfo", "-Xcompiler", "-fPIC", "-shared",
]


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh")) +
                  [os.path.join(ROOT, "include", "playsnark_b200.h")])


def stale(target: str, sources) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    srcs = _sources()
    if not force and not stale(LIB, srcs):
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["--threads", "2", "-o", LIB] + CU_SOURCES
    print("[playsnark_b200] " + " ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd, cwd=ROOT)
    return LIB


def build_host_emulation(out_dir: str) -> str:
    """TEST-ONLY: the same sources with -DPS_HOST_EMU (kernel bodies driven by serial loops)."""
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libps_hostemu.so")
    if stale(so, _sources()):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-x", "c++", "-DPS_HOST_EMU", "-shared", "-fPIC", "-o", so] + CU_SOURCES,
                              cwd=ROOT)
    return so


if __name__ == "__main__":
    build_cuda(force="--force" in sys.argv, verbose="-v" in sys.argv)
