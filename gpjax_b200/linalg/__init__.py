"""gpjax.linalg mirror (gpjax/linalg/__init__.py:23-37)."""
from .operations import diag, logdet, lower_cholesky, solve
from .operators import Dense, Diagonal, Identity, LinearOperator, Triangular
from .utils import PSD, add_jitter, psd

__all__ = ["LinearOperator", "Dense", "Diagonal", "Identity", "Triangular", "lower_cholesky", "solve", "logdet", "diag",
           "psd", "PSD", "add_jitter"]
