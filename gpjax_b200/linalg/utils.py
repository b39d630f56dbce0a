"""gpjax/linalg/utils.py:21-65."""
from __future__ import annotations

import torch


class _PSD:
    def __repr__(self):
        return "PSD"


PSD = _PSD()


def psd(A):
    """Attach the PSD marker (purely an annotation, utils.py:21-36)."""
    A.annotations = set(getattr(A, "annotations", set())) | {PSD}
    return A


def add_jitter(matrix: torch.Tensor, jitter=1e-6) -> torch.Tensor:
    if matrix.ndim != 2:
        raise ValueError(f"Expected 2D matrix, got {matrix.ndim}D array")
    if matrix.shape[0] != matrix.shape[1]:
        raise ValueError(f"Expected square matrix, got shape {tuple(matrix.shape)}")
    out = matrix.clone()
    out.diagonal().add_(jitter)
    return out
