"""Build libmmgpu.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

Five translation units (host ABI + four kernel files) are compiled in parallel to objects under csrc/_obj/ and
linked into one shared library; only units whose sources changed are recompiled."""
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
HDRS = ["mm_common.cuh", "mm_map.cuh", "mm_libm.cuh", "mm_tan_exceptions.h", "mm_sbpos.cuh", "mm_fir.cuh", "mm_htaps.h"]
# unit -> headers it depends on (besides include/mmgpu.h)
UNITS = {
    "mmgpu.cu": ["mm_common.cuh", "mm_map.cuh", "mm_libm.cuh", "mm_fir.cuh", "mm_htaps.h"],
    "mm_kernels_map.cu": ["mm_common.cuh", "mm_map.cuh", "mm_libm.cuh", "mm_tan_exceptions.h", "mm_sbpos.cuh"],
    "mm_kernels_fir.cu": ["mm_common.cuh", "mm_map.cuh", "mm_libm.cuh", "mm_fir.cuh", "mm_htaps.h"],
    "mm_kernels_mc.cu": HDRS,
    "mm_kernels_dmvr.cu": ["mm_common.cuh"],
}
ABI_H = os.path.join(os.path.dirname(HERE), "include", "mmgpu.h")
OUT = os.path.join(HERE, "libmmgpu.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    # parity: no FMA contraction, IEEE division/sqrt, denormals kept (also enforced by explicit *_rn intrinsics)
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-O2",
]


def library_path():
    """MMGPU_LIB overrides the library (kernel-variant experiments); default: the in-tree build."""
    return os.environ.get("MMGPU_LIB", OUT)


def _obj(unit):
    return os.path.join(OBJ, unit[:-3] + ".o")


def _stale(unit):
    o = _obj(unit)
    if not os.path.exists(o):
        return True
    t = os.path.getmtime(o)
    deps = [os.path.join(CSRC, unit), ABI_H] + [os.path.join(CSRC, h) for h in UNITS[unit]]
    return any(os.path.getmtime(d) > t for d in deps)


def needs_build():
    if not os.path.exists(OUT):
        return True
    return any(_stale(u) or os.path.getmtime(_obj(u)) > os.path.getmtime(OUT) for u in UNITS)


def build_library(force=False, verbose=False, out=None, extra=None):
    """out/extra: build a variant library (other nvcc flags) somewhere else; objects then go to <out>.obj/."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = list(extra or []) + os.environ.get("MMGPU_NVCC_EXTRA", "").split()
    variant = out is not None
    target = out or OUT
    objdir = (target + ".obj") if variant else OBJ
    if not force and not variant and not needs_build():
        return OUT
    os.makedirs(objdir, exist_ok=True)

    def compile_unit(unit):
        o = os.path.join(objdir, unit[:-3] + ".o")
        if not force and not variant and not _stale(unit):
            return o
        cmd = [nvcc] + NVCC_FLAGS + extra + ["-c", "-o", o, os.path.join(CSRC, unit)]
        if verbose:
            cmd += ["-Xptxas", "-v"]
        subprocess.check_call(cmd)
        return o

    with ThreadPoolExecutor(max_workers=len(UNITS)) as ex:
        objs = list(ex.map(compile_unit, UNITS))
    subprocess.check_call([nvcc, "-shared", "-o", target] + objs)
    return target
