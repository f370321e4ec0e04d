"""
Builds libsdft_b200.so IN-TREE with nvcc for sm_100a (cross-compiles without a GPU).

    python -m sdft_b200.build [--force]

The shared object sits next to this file (sdft_b200/libsdft_b200.so); it is git-ignored but travels to the
GPU box with the repository snapshot.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsdft_b200.so")
SOURCES = [os.path.join(CSRC, "sdft_b200.cu")]
DEPENDS = SOURCES + [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".hpp"))] + [
    os.path.join(HERE, "..", "include", "sdft_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-split-compile", "0",          # ptxas over the ~120 kernel variants in parallel (same code, a third of the time)
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden",
    "-shared",
]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPENDS if os.path.exists(d))


def build(force=False, verbose=False, out=None, extra=()):
    """Builds the library.  `out`/`extra` build an experimental variant (other file name, extra nvcc
    flags such as -DSDFT_B200_MINBLOCKS=5) that is selected at run time with SDFT_B200_LIB=<path>."""
    target = out or LIB
    if out is None and not force and not stale():
        return LIB
    nvcc = find_nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build libsdft_b200.so (there is no CPU fallback)")
    cmd = [nvcc] + NVCC_FLAGS + list(extra) + (["-Xptxas", "-v"] if verbose else []) + ["-o", target] + SOURCES
    subprocess.run(cmd, check=True)
    return target


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
