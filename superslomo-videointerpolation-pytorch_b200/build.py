"""Builds libssm_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python superslomo-videointerpolation-pytorch_b200/build.py [--force] [--verbose]
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libssm_b200.so")
SOURCES = ["ssm_abi.cu"]


def _deps():
    """every file the library is compiled from: csrc/*.cu, csrc/*.cuh (new headers are picked up without
    editing a list) and the public header"""
    import glob
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh"))) \
        + [os.path.join(ROOT, "include", "ssm_b200.h")]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libssm_b200.so cannot be built (there is no CPU fallback)")


def needs_build():
    if not os.path.exists(LIB):
        return True
    built = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > built for d in _deps())


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    cmd = [
        _nvcc(),
        "-gencode", "arch=compute_100a,code=sm_100a",
        "-O3", "-lineinfo", "-std=c++17",
        "-Xcompiler", "-fPIC,-ffp-contract=off",
        "-I", os.path.join(ROOT, "include"),
        "-shared", "-o", LIB,
    ] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
