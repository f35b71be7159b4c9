"""Build recipe of the CUDA library (in-tree, sm_100a only).

    python -m femtech_b200.build          # builds femtech_b200/libftb200.so

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels with
gpurun snapshots.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libftb200.so")
SOURCES = [os.path.join(CSRC, "ftb200_capi.cu")]
DEPS = SOURCES + [os.path.join(CSRC, "ftb200_kernels.cuh"), os.path.join(CSRC, "ftb200_brick.cuh"), os.path.join(CSRC, "hex8_element.cuh"),
                  os.path.join(os.path.dirname(HERE), "include", "ftb200.h")]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared"]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False):
    """Compile libftb200.so if missing or older than its sources."""
    if not force and not stale():
        return LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        sys.stderr.write(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
