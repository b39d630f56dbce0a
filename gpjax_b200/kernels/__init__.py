from .base import AbstractKernel
from .computations import AbstractKernelComputation, DenseKernelComputation
from .stationary import RBF, Matern12, Matern32, Matern52, StationaryKernel

__all__ = ["AbstractKernel", "StationaryKernel", "RBF", "Matern12", "Matern32", "Matern52", "AbstractKernelComputation",
           "DenseKernelComputation"]
