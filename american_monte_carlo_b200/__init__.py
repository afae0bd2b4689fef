"""B200-native Longstaff-Schwartz hot path (GBM path simulation + LSM backward induction).

Host-side mirror of /root/reference/american_monte_carlo.py:72-197 over the C ABI of include/amc.h.
Importing the package does not need a GPU; the first call that computes anything does, and raises without one.
"""
from .api import (apply_exercise, estimate_continuation_values, main, perform_backward_iteration,  # noqa: F401
                  Context, ContinuationValues, DevicePaths, LsmResult, compute_ccr_exposures, default_context,
                  generate_asset_paths,  # noqa: F401
                  get_basis_polynomials, intrinsic_value, lsm_price, lsm_price_batch, lsmc_option_pricing, paths_from_host,
                  paths_from_normals, precompute_barrier_hit_matrix, regression_estimate, set_default_context,
                  shard_range)

__all__ = ["apply_exercise", "estimate_continuation_values", "main", "perform_backward_iteration", "Context", "ContinuationValues", "DevicePaths", "LsmResult", "compute_ccr_exposures", "default_context", "generate_asset_paths",
           "get_basis_polynomials", "intrinsic_value", "lsm_price", "lsm_price_batch", "lsmc_option_pricing", "paths_from_host",
           "paths_from_normals", "precompute_barrier_hit_matrix", "regression_estimate", "set_default_context",
           "shard_range"]
