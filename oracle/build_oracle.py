"""TEST INFRASTRUCTURE: compiles oracle/rwrt_oracle.c -> oracle/_build/librwrt_oracle.so with gcc.

No fp contraction, no fast-math: the reference evaluates every expression as separate IEEE
double operations (NumPy / numba without fastmath), and so must its restatement.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "rwrt_oracle.c")
DEPS = [SRC, os.path.join(HERE, "zgeev3.inc")]
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "librwrt_oracle.so")


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS):
        return OUT
    cmd = ["gcc", "-O2", "-std=gnu11", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
           "-Wall", "-Wno-unused-function", "-o", OUT, SRC, "-lm"]
    subprocess.run(cmd, check=True, cwd=HERE)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
