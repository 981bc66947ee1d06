"""Builds libaz_b200.so (the C-ABI engine) in-tree with nvcc for sm_100a.

    python -m alphazero_openspiel_b200.build          # or __graft_entry__.build()

-fmad=false: the PUCT / backup arithmetic must round every fp64 operation like the reference's Python
floats; the kernels additionally use explicit __d*_rn intrinsics.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libaz_b200.so")
SOURCES = [os.path.join(HERE, "csrc", "az_engine.cu"), os.path.join(HERE, "csrc", "az_resnet.cu")]
DEPS = SOURCES + [os.path.join(HERE, "csrc", "az_games.cuh"),
                  os.path.join(HERE, "..", "include", "az_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--fmad=false", "-Xcompiler", "-fPIC", "-shared"]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout)
    if verbose:
        print(res.stdout)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
