"""gpjax_b200 -- B200-native drop-in for GPJax's data-parallel inference hot path.

    import gpjax_b200 as gpx
    kernel = gpx.kernels.RBF()                       # .gram() / .cross_covariance() -> fused sm_100a Gram tiles
    prior = gpx.gps.Prior(mean_function=gpx.mean_functions.Zero(), kernel=kernel)
    posterior = prior * gpx.likelihoods.Gaussian(num_datapoints=D.n)
    gpx.objectives.conjugate_mll(posterior, D)       # blocked DMMA Cholesky + analytic backward
    gpx.fit(model=posterior, objective=lambda p, d: -gpx.objectives.conjugate_mll(p, d), train_data=D,
            optim=gpx.optim.adam(1e-2))

Only the path named in DESIGN.md is implemented; everything computes on CUDA (no CPU fallback).
"""
from . import (dataset, distributions, fit as _fit_mod, gps, kernels, likelihoods, linalg, mean_functions, objectives,
               optim, parameters, variational_families)
from .dataset import Dataset
from .fit import fit, fit_lbfgs, fit_scipy, get_batch

__version__ = "0.1.0"
__all__ = ["Dataset", "fit", "fit_scipy", "fit_lbfgs", "get_batch", "kernels", "linalg", "objectives", "gps", "likelihoods",
           "mean_functions", "variational_families", "parameters", "optim", "distributions", "dataset"]
