"""Build libmmgpu.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "mmgpu.cu")
DEPS = [os.path.join(HERE, "csrc", f) for f in ("mmgpu.cu", "mm_kernels.cuh", "mm_map.cuh", "mm_libm.cuh")] + \
       [os.path.join(os.path.dirname(HERE), "include", "mmgpu.h")]
OUT = os.path.join(HERE, "libmmgpu.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    # parity: no FMA contraction, IEEE division/sqrt, denormals kept (also enforced by explicit *_rn intrinsics)
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-O2", "-shared",
]


def library_path():
    """MMGPU_LIB overrides the library (kernel-variant experiments); default: the in-tree build."""
    return os.environ.get("MMGPU_LIB", OUT)


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build_library(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + os.environ.get("MMGPU_NVCC_EXTRA", "").split() + ["-o", OUT, SRC]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    subprocess.check_call(cmd)
    return OUT
