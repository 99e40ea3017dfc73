"""Builds librwrt.so (the CUDA engine behind include/rwrt.h) in-tree with nvcc for sm_100a.

    python rossby-wave-ray-tracing_b200/build.py [--force]

nvcc cross-compiles without a GPU.  Three floating-point regimes, per translation unit:
  * rwrt_bs.cu, rwrt_init.cu, rwrt_grid.cu, rwrt_rk45_strict.cu : -fmad=false (reference operation order,
    bit-exact fields / roots / cell indices; the strict integrator)
  * rwrt_rk45_fast.cu, rwrt_api.cu                              : default contraction (FMA)
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "librwrt.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--use_fast_math=false"]
UNITS = {
    "rwrt_api.cu": [],
    "rwrt_bs.cu": ["-fmad=false"],
    "rwrt_init.cu": ["-fmad=false"],
    "rwrt_grid.cu": ["-fmad=false"],
    "rwrt_rk45_fast.cu": [],
    "rwrt_rk45_strict.cu": ["-fmad=false"],
    "rwrt_shsf.cu": [],
    "rwrt_hx.cpp": [],       # host code only (the host half of rwrt_real2d_host), straight to the host compiler
}


def _deps():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "rwrt.h"),
                                                                os.path.abspath(__file__)]


def _compile(unit, flags, log, objdir=None):
    obj = os.path.join(objdir or OBJ, os.path.splitext(unit)[0] + ".o")
    cmd = [NVCC, *ARCH, *[c for c in COMMON if not c.startswith("--use_fast_math")], *flags, "-c",
           os.path.join(CSRC, unit), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log.append(f"$ {' '.join(cmd)}\n{r.stdout}{r.stderr}")
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {unit}:\n{r.stdout}\n{r.stderr}")
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in _deps()):
        return LIB
    log = []
    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(lambda kv: _compile(kv[0], kv[1], log), UNITS.items()))
    cmd = [NVCC, *ARCH, "-shared", "-o", LIB, *objs, "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log.append(f"$ {' '.join(cmd)}\n{r.stdout}{r.stderr}")
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(os.path.join(OBJ, "build.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return LIB


def build_variant(out: str, extra: list) -> str:
    """A/B build of the same ABI with extra nvcc flags (e.g. -DRWRT_TMEM_KS=0) into `out`; load it with RWRT_LIB=out."""
    objdir = out + ".obj"
    os.makedirs(objdir, exist_ok=True)
    log = []
    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(lambda kv: _compile(kv[0], kv[1] + extra, log, objdir), UNITS.items()))
    r = subprocess.run([NVCC, *ARCH, "-shared", "-o", out, *objs, "-lcudart"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(os.path.join(objdir, "build.log"), "w") as f:
        f.write("\n".join(log))
    return out


if __name__ == "__main__":
    if "--variant" in sys.argv:     # python build.py --variant out.so -DFOO=1 ...
        i = sys.argv.index("--variant")
        print(build_variant(os.path.abspath(sys.argv[i + 1]), sys.argv[i + 2:]))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
