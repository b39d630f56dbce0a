"""ctypes prototypes for the C ABI declared in ``include/gpjax_b200.h``.

``declare(lib)`` attaches argtypes/restype to every exported entry point of a loaded
``ctypes.CDLL``.  The table below is the single Python-side statement of the ABI; the CPU
test-suite also checks it against the header and against the symbols the shared library exports.
"""
from __future__ import annotations

import ctypes as C

vp = C.c_void_p
i64 = C.c_int64
i32 = C.c_int
f64 = C.c_double

# name -> (restype, [argtypes])
PROTOTYPES = {
    "gpb_version": (C.c_char_p, []),
    "gpb_max_input_dim": (i32, []),
    "gpb_block_size": (i64, []),
    "gpb_block_size_for": (i64, [i64]),
    "gpb_profile_reset": (None, [i32]),
    "gpb_debug_set_gemm_variant": (None, [i32]),
    "gpb_profile_read": (i32, [vp, vp, vp]),
    "gpb_profile_read_ozaki": (i32, [vp, vp, vp]),
    "gpb_gram": (i32, [vp, i32, i64, i64, i32, vp, i64, vp, i64, vp, i32, vp, f64, vp, i32, vp, i64]),
    "gpb_gram_bwd_workspace_bytes": (i64, [i64, i64, i32]),
    "gpb_gram_bwd": (
        i32,
        [vp, i32, i64, i64, i32, vp, i64, vp, i64, vp, i32, vp, vp, i64, f64, vp, i64, vp, vp, vp, i64, vp, i64],
    ),
    "gpb_factor_workspace_bytes": (i64, [i64, i32, i32]),
    "gpb_potrf_lower": (i32, [vp, i64, vp, i64, i32, vp, i64, i64, i32, i32, vp]),
    "gpb_diag_inverses": (i32, [vp, i64, vp, i64, vp, i64, i64, i32, i32]),
    "gpb_trsv_lower": (i32, [vp, i64, vp, i64, i32, vp, vp, i64, i64, i32, i32]),
    "gpb_trsm_lower_left": (i32, [vp, i64, i64, vp, i64, i32, vp, i64, vp, i64, i64, i32, i32]),
    "gpb_sum_log_diag": (i32, [vp, i64, vp, i64, vp]),
    "gpb_potri_lower": (i32, [vp, i64, vp, i64, vp, i64, vp, i64, i64, i32, i32]),
    "gpb_gemm": (i32, [vp, i64, i64, i64, f64, vp, i64, i32, vp, i64, i32, f64, vp, i64, i32]),
    "gpb_ozaki_available": (i32, []),
    "gpb_set_ozaki_slices": (None, [i32]),
    "gpb_get_ozaki_slices": (i32, []),
    "gpb_ozaki_auto_planes": (i32, [i64, f64, f64, f64]),
    "gpb_ozaki_slice": (i32, [vp, i64, i64, vp, i64, i32, vp, i64, vp]),
    "gpb_ozaki_gemm": (i32, [vp, i64, i64, i64, i32, vp, i64, vp, vp, i64, vp, f64, vp, i64, i32]),
    "gpb_igemm_i8": (i32, [vp, i64, i64, i64, vp, i64, vp, i64, vp, i64]),
    "gpb_mll_workspace_bytes": (i64, [i64, i32]),
    "gpb_mll_forward": (
        i32,
        [vp, i32, i64, i32, vp, i64, vp, vp, i32, vp, vp, vp, f64, vp, i64, vp, i64, vp, vp, vp],
    ),
    "gpb_mll_backward": (
        i32,
        [vp, i32, i64, i32, vp, i64, vp, i32, vp, vp, vp, i64, vp, i64, vp, vp, vp, vp, vp, vp],
    ),
    "gpb_sgpr_workspace_bytes": (i64, [i64, i32, i64]),
    "gpb_sgpr_stats_count": (i64, [i64]),
    "gpb_sgpr_stats": (
        i32,
        [vp, i32, i64, i64, i32, vp, i64, vp, vp, i64, vp, i32, vp, vp, vp, f64, i64, vp, i64, vp],
    ),
    "gpb_sgpr_stats_raw": (
        i32,
        [vp, i32, i64, i64, i32, vp, i64, vp, vp, i64, vp, i32, vp, vp, vp, f64, i64, vp, i64, vp],
    ),
    "gpb_sgpr_finish": (i32, [vp, i32, i64, i32, vp, i64, vp, i32, vp, vp, i64, vp, i64, vp, i32, vp, vp]),
    "gpb_sgpr_grad_local": (
        i32,
        [vp, i32, i64, i64, i32, vp, i64, vp, vp, i64, vp, i32, vp, vp, vp, i64, vp, i64, vp, vp, vp],
    ),
    "gpb_sgpr_grad_finish": (
        i32,
        [vp, i32, i64, i32, vp, i64, vp, i32, vp, vp, i64, vp, i64, vp, vp, vp, vp, vp, vp],
    ),
    "gpb_svgp_finish": (
        i32,
        [vp, i32, i64, i32, vp, i64, vp, i32, vp, vp, vp, vp, vp, i64, f64, f64, i64, vp, i64, vp, i32, vp, vp],
    ),
    "gpb_svgp_grad_finish": (
        i32,
        [vp, i32, i64, i32, vp, i64, vp, i32, vp, vp, f64, i64, vp, i64, vp, vp, i64, vp, vp, vp, vp, vp, vp, vp, i64],
    ),
    "gpb_nccl_version": (i32, []),
    "gpb_nccl_unique_id": (i32, [vp]),
    "gpb_nccl_comm_init_rank": (i32, [vp, i32, vp, i32]),
    "gpb_nccl_comm_destroy": (i32, [vp]),
    "gpb_allreduce_f64": (i32, [vp, vp, vp, i64]),
}

ERRORS = {
    -1: "GPB_ERR_INVALID (bad argument)",
    -2: "GPB_ERR_UNSUPPORTED (capability not compiled in, e.g. input dimension too large)",
    -3: "GPB_ERR_LAUNCH (CUDA launch failed)",
    -4: "GPB_ERR_WORKSPACE (workspace too small)",
}


def declare(lib: C.CDLL) -> C.CDLL:
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed: {ERRORS.get(rc, rc)}")
