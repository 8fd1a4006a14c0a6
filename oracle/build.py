"""Compile the oracle's C restatement (oracle/raytrace.c) into oracle/_build/.

Test infrastructure: building the checker is not using it.  The reference itself is
pure Python (0 compiled sources), so there is no ``oracle/_ref`` to build.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "liboracle_rt.so")


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "raytrace.c")
    if (not force) and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(src):
        return LIB
    os.makedirs(OUT, exist_ok=True)
    cmd = ["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-shared", "-fPIC", "-o", LIB, src, "-lm"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
