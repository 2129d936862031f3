"""mcmcglm_b200 -- B200-native CGGibbs engine behind the mcmcglm() API of mathiaslj/mcmcglm.

Only the slice-within-Gibbs coordinate-update path runs here (CUDA, sm_100a, via libcggibbs.so);
see DESIGN.md.  Importing the package does not load CUDA; constructing an Engine does.
"""
from ._lib import CggError  # noqa: F401
from .engine import Engine  # noqa: F401
from .api import (mcmcglm, samples, coef, quantile, log_potential_from_betaj, update_linear_predictor,  # noqa: F401
                  mcmcglm_across_tuningparams, slice_stepping_out, dist_normal, dist_laplace, dist_student_t,
                  dist_gamma, dist_exponential, gaussian, binomial, poisson, negative_binomial, check_family,
                  extract_model_data, McmcGlm, compare_eta_comptime, compare_eta_comptime_across_nvars,
                  generate_normal_data)

__all__ = ["Engine", "CggError", "mcmcglm", "samples", "coef", "quantile", "log_potential_from_betaj",
           "update_linear_predictor", "mcmcglm_across_tuningparams", "slice_stepping_out", "dist_normal",
           "dist_laplace", "dist_student_t", "dist_gamma", "dist_exponential", "gaussian", "binomial", "poisson",
           "negative_binomial", "compare_eta_comptime", "compare_eta_comptime_across_nvars", "generate_normal_data"]
