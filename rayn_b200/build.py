"""In-tree build of librayn_b200.so for sm_100a (nvcc cross-compiles without a GPU).

The arithmetic flags are part of the parity contract (csrc/detmath.h): no implicit FMA
contraction, IEEE division and square root, no flush-to-zero.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "librayn_b200.so")
SOURCES = [os.path.join(CSRC, "api.cu"), os.path.join(CSRC, "host_inputs.cpp")]
DEPS = SOURCES + [os.path.join(CSRC, f) for f in ("rt_kernels.cuh", "rt_device.cuh", "detmath.h")] + [
    os.path.join(HERE, "..", "include", "rayn_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "--fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math", "-shared",
]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        if not os.path.exists(os.path.join(OUT_DIR, "rayn_host")):
            build_host()
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + SOURCES
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building librayn_b200.so")
    build_host()
    return OUT


def build_host():
    """C++ host stand-in (rayn_b200/host): links against the C ABI only."""
    r = subprocess.run(["make", "-C", os.path.join(HERE, "host"), "-B"], capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("building rayn_host failed")
    return os.path.join(OUT_DIR, "rayn_host")


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(OUT)
