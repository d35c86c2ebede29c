"""Builds the C part of the oracle (test infrastructure) into oracle/_build/libps_oracle.so.

The reference itself (Go, with un-vendored arithmetic modules and no Go toolchain in this image)
cannot be compiled, so there is no oracle/_ref/ build; see DESIGN.md."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "ps_prover.c")      # includes ps_oracle.c: one translation unit
DEPS = [SRC, os.path.join(HERE, "ps_oracle.c")]
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libps_oracle.so")


def build() -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in DEPS):
        subprocess.check_call(["gcc", "-O2", "-std=gnu11", "-fopenmp", "-shared", "-fPIC", "-o", LIB, SRC])
    return LIB


if __name__ == "__main__":
    print(build())
