from .base import AbstractKernel, CombinationKernel, Constant, ProductKernel, SumKernel
from .computations import AbstractKernelComputation, ConstantDiagonalKernelComputation, DenseKernelComputation
from .stationary import (RBF, Matern12, Matern32, Matern52, Periodic, PoweredExponential, RationalQuadratic,
                         StationaryKernel, White)

__all__ = ["AbstractKernel", "StationaryKernel", "RBF", "Matern12", "Matern32", "Matern52", "RationalQuadratic",
           "PoweredExponential", "Periodic", "White", "Constant", "CombinationKernel", "SumKernel", "ProductKernel",
           "AbstractKernelComputation", "DenseKernelComputation", "ConstantDiagonalKernelComputation"]
