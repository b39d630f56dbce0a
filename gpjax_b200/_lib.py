"""Loader for the CUDA library.  There is NO CPU fallback: a missing library is a hard error."""
from __future__ import annotations

import ctypes
import os

from . import _abi

_LIB = None
# GPB_LIB_PATH: measurement hook (a build with another block size, GPB_NB=... scripts/nb_sweep.py); never a fallback
LIB_PATH = os.environ.get("GPB_LIB_PATH") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libgpjax_b200.so")


class ExtensionMissingError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    """The loaded ``libgpjax_b200.so`` (built by ``python -m gpjax_b200.build``)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ExtensionMissingError(
                f"{LIB_PATH} not found. gpjax_b200 has no CPU/PyTorch fallback: build the sm_100a "
                "CUDA library first with `python -m gpjax_b200.build` (needs nvcc)."
            )
        _LIB = _abi.declare(ctypes.CDLL(LIB_PATH))
        if os.environ.get("GPB_GEMM_VARIANT"):  # kernel-tuning hook (scripts/gemm_bench.py, profiles/)
            _LIB.gpb_debug_set_gemm_variant(int(os.environ["GPB_GEMM_VARIANT"]))
    return _LIB


def require_cuda(*tensors) -> None:
    import torch

    for t in tensors:
        if t is None:
            continue
        if not isinstance(t, torch.Tensor) or not t.is_cuda:
            raise RuntimeError(
                "gpjax_b200 computes on CUDA devices only (no CPU fallback); got a "
                f"{'CPU tensor' if isinstance(t, torch.Tensor) else type(t).__name__}"
            )
        if t.dtype != torch.float64:
            raise TypeError(f"gpjax_b200 is float64-only (GPJax runs with x64); got {t.dtype}")
        cur = torch.cuda.current_device()
        if t.device.index != cur:  # launches go to the CURRENT device's stream: refuse instead of using the wrong one
            raise RuntimeError(f"tensor lives on cuda:{t.device.index} but the current device is cuda:{cur}; wrap the call "
                               "in `with torch.cuda.device(tensor.device):`")
