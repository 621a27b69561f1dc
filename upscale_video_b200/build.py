"""Builds libb2sr.so (the C-ABI library of include/b2sr.h) in-tree with nvcc for sm_100a.

    python -m upscale_video_b200.build [--force]

The shared object is written next to this file so that it travels with the repo snapshot to the GPU box and is
the file the Python shim loads (upscale_video_b200/engine.py).  cudart is linked statically and the driver API
(cuTensorMapEncodeTiled) is resolved at run time, so the library loads on a machine without a GPU driver.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libb2sr.so")
SOURCES = ["engine.cu"]
HEADERS = ["common.cuh", "tc_conv.cuh", "tc_gconv.cuh", "graph_exec.cuh", "simple_kernels.cuh", "nlm.cuh", "nlm_host.inl", os.path.join("..", "..", "include", "b2sr.h")]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build_lib(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-cudart", "static",
           "-Xptxas", "-v" if verbose else "-O3",
           "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv))
