"""Build libwann_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m wann_b200.build            # build if stale
    python -m wann_b200.build --force
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libwann_b200.so")
SOURCES = ["wann_api.cu"]
DEPS = ["wann_api.cu", "conv_stack_tc.cuh", "conv_stack_split.cuh", "aux_kernels.cuh", "ptx.cuh", "array_tree.hpp", "gptq.hpp",
        os.path.join("..", "..", "include", "wann_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fopenmp", "-shared", "-lgomp"]


def stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force=False, verbose=False):
    if not force and not stale():
        return OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + SOURCES
    env = dict(os.environ)
    # the image exports CC/CXX=/opt/gcc/...; let nvcc pick its default host compiler
    subprocess.check_call(cmd, cwd=CSRC, env=env)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
