"""In-tree build of libdmsa_b200.so (hand-written CUDA, sm_100a only)."""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "csrc", "dmsa_b200.cu")
HOST_SRC = os.path.join(_HERE, "csrc", "host_solve.cpp")
HOST_OBJ = os.path.join(_HERE, "lib", "host_solve.o")
DEPS = [os.path.join(_HERE, "csrc", f) for f in ("host_solve.cpp", "dmsa_b200.cu", "kernels_cost.cuh", "kernels_pose.cuh", "kernels_sets.cuh", "kernels_solve.cuh", "kernels_sort.cuh", "kernels_chol.cuh", "kernels_knn.cuh", "kernels_pre.cuh", "dmsa_b200_pre.inl", "dmsa_b200_io.inl", "se3_math.cuh", "pdl.cuh")] + [
    os.path.join(os.path.dirname(_HERE), "include", "dmsa_b200.h")]
OUT = os.path.join(_HERE, "lib", "libdmsa_b200.so")
OUT_FMA = os.path.join(_HERE, "lib", "libdmsa_b200_fma.so")  # experiment: the same kernels with FMA contraction allowed (scripts/fma_deviation.py)

# -fmad=false: the reference's float arithmetic has no FMA contraction (CMakeLists.txt:13-17, baseline x86-64);
# the kernels use explicit fma() where fusion is wanted (J^T J) and explicit *_rn intrinsics on the parity-critical path.
NVCC_FLAGS = (["-DDMSA_TIMELINE=" + os.environ["DMSA_TIMELINE"]] if os.environ.get("DMSA_TIMELINE") else []) + ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false", "-Xcompiler", "-fPIC", "-shared",
              # host side (LM solve): AVX2 vector loops, still no FMA contraction so results equal the scalar sequence
              "-Xcompiler", "-O3,-mavx2,-ffp-contract=off"]


def nvcc_path():
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def is_stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build_fma_variant():
    """The experiment library of scripts/fma_deviation.py: -fmad=true and plain float operators on the cost path."""
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-std=c++17", "-O3", "-mavx2", "-ffp-contract=off", "-fPIC", "-pthread", "-c", "-o", HOST_OBJ, HOST_SRC])
    flags = [f for f in NVCC_FLAGS if f != "-fmad=false"] + ["-fmad=true", "-DDMSA_ALLOW_FMA"]
    cmd = [nvcc_path()] + (["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []) + flags + ["-o", OUT_FMA, SRC, HOST_OBJ]
    subprocess.check_call(cmd)
    return OUT_FMA


def build_library(force=False, verbose=False):
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    if not force and not is_stale():
        return OUT
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-std=c++17", "-O3", "-mavx2", "-ffp-contract=off", "-fPIC", "-pthread", "-c", "-o", HOST_OBJ, HOST_SRC])
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT, SRC, HOST_OBJ]
    env = dict(os.environ)
    # the image exports CC/CXX=/opt/gcc/bin/*; nvcc must use the distro host compiler
    if os.path.exists("/usr/bin/g++"):
        cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
    subprocess.check_call(cmd, env=env)
    return OUT


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
