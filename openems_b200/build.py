"""Builds libopenems_b200.so in-tree with nvcc for sm_100a (no GPU needed to compile)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "lib", "libopenems_b200.so")
SOURCES = ["engine.cu", "host/synthetic_operator.cpp"]
DEPS = ["kernels.cuh", "kernels_fused.cuh", "kernels_fused_tma.cuh", "kernels_xslab.cuh", "kernels_xslab_tma.cuh", "engine.h", "abi.inc", "entry_set.h", "host/synthetic_operator.h", "../../include/openems_b200.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    # parity contract: no FMA contraction, flush denormals, IEEE div/sqrt
    "-fmad=false", "-ftz=true", "-prec-div=true", "-prec-sqrt=true",
    # hidden visibility + -Bsymbolic: only the extern "C" API is exported and the library's own C++ symbols (class
    # Engine ...) can neither clash with nor be interposed by the host application's (openEMS has an Engine too)
    "-Xcompiler", "-fPIC,-fopenmp,-O2,-ffp-contract=off,-fvisibility=hidden",
    "-shared", "-lgomp", "-Xlinker", "-Bsymbolic",
]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    files = [os.path.join(SRC, s) for s in SOURCES + DEPS] + [os.path.abspath(__file__)]
    return any(os.path.exists(f) and os.path.getmtime(f) > t for f in files)


def build(force=False, verbose=False, defines=(), out=None):
    """defines/out: tuning variants for sweeps (tools/sweep.py); the product build uses neither"""
    global OUT
    if out is not None:
        saved, OUT = OUT, out
        try:
            return _build(True, verbose, defines)
        finally:
            OUT = saved
    if not force and not needs_build():
        return OUT
    return _build(force, verbose, defines)


def _build(force, verbose, defines):
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    srcs = [os.path.join(SRC, s) for s in SOURCES if os.path.exists(os.path.join(SRC, s))]
    cmd = [nvcc] + NVCC_FLAGS + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else [])
    if os.path.exists("/usr/bin/g++"):
        cmd += ["-ccbin", "/usr/bin/g++"]
    cmd += ["-I", os.path.join(HERE, "..", "include"), "-o", OUT] + srcs
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout)
        raise RuntimeError("nvcc failed building libopenems_b200.so")
    if verbose:
        print(res.stdout)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
