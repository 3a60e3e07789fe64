"""vvc-extension-mm_b200 -- B200 (sm_100a) multi-model 360 MC + block-matching distortion path.

The product is the C-ABI shared library ``libmmgpu.so`` (include/mmgpu.h, csrc/); this package is
the thin Python host mirror used by tests, bench.py and __graft_entry__.py.  Import it as
``vvc_extension_mm_b200`` (shim module at the repo root; the directory name carries a hyphen).
There is no CPU fallback: every compute call raises MmGpuError when the CUDA library or a
sm_100 device is missing.
"""
from .build import build_library, library_path  # noqa: F401
from .api import (  # noqa: F401
    MmGpu, MmGpuError, MmPlan, Block, MODEL_NAMES,
    CLASSIC, MPA_FB, MPA_LR, MPA_TB, TAN, T3D, ROT, GED_X, GED_Y, GED_Z, GED_CAMPOSE,
    OOB_PRED, OOB_ME, DIST_SAD, DIST_HAD, HAD_SCALAR, HAD_SIMD,
    rodrigues, fixed_to_float, exported_symbols, load_library, dmvr_decide, DMVR_DTYPE, JOB_DTYPE,
)
