"""Builds libpsinfer.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

    python -m partapp_b200.build [--force]

nvcc cross-compiles for sm_100a without a GPU.  -fmad=false / -ffp-contract=off keep every fp32/fp64
multiply and add separately rounded (DESIGN.md "Arithmetic contract"); -lineinfo lets ncu map SASS to source.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "csrc", "psinfer.cu")
DEPS = [SRC, os.path.join(HERE, "csrc", "ps_kernels.cuh"), os.path.join(HERE, "csrc", "ps_geometry.hpp"),
        os.path.join(ROOT, "include", "psinfer.h")]
OUT = os.path.join(HERE, "libpsinfer.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-O2",
    "-shared",
]


def nvcc_path():
    for p in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if p and (os.path.isabs(p) and os.path.exists(p) or not os.path.isabs(p)):
            return p
    return "nvcc"


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT, SRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libpsinfer.so")
    if verbose:
        sys.stderr.write(res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
