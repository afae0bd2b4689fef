"""Drop-in module: the reference's import name on top of the B200 CUDA path.

`from american_monte_carlo import lsmc_option_pricing, get_quantlib_option, generate_asset_paths, intrinsic_value`
(unit_test.py:3 of the reference) resolves here when this repository precedes the reference on sys.path.  The hot
path (amc.py:72-197) runs in libamc.so on the GPU; `get_quantlib_option` is a QuantLib-free benchmark stand-in
(american_monte_carlo_b200/benchmarks.py); plotting and the notebook driver are out of scope.
"""
from american_monte_carlo_b200.api import (apply_exercise, compute_ccr_exposures, estimate_continuation_values,  # noqa: F401
                                           main, perform_backward_iteration, generate_asset_paths, get_basis_polynomials, intrinsic_value,  # noqa: F401
                                           lsmc_option_pricing, precompute_barrier_hit_matrix, regression_estimate)
from american_monte_carlo_b200.benchmarks import get_quantlib_option  # noqa: F401

__all__ = ["apply_exercise", "compute_ccr_exposures", "estimate_continuation_values", "main",
           "perform_backward_iteration", "generate_asset_paths", "get_basis_polynomials", "intrinsic_value", "lsmc_option_pricing",
           "precompute_barrier_hit_matrix", "regression_estimate", "get_quantlib_option"]
