"""ctypes binding of libamc.so (include/amc.h).  No fallback: a missing library or GPU raises."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AMC_LIBAMC") or os.path.join(HERE, "libamc.so")    # override: A/B of two builds

AMC_MAX_K = 11
F64, F32 = 0, 1
BASIS_ID = {"Power": 0, "Chebyshev": 1, "Legendre": 2, "Laguerre": 3}
ERR_VALUE = 1

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)


class LsmSpec(C.Structure):
    _fields_ = [("K", C.c_double), ("r", C.c_double), ("dt", C.c_double), ("barrier", C.c_double),
                ("scaling_factor", C.c_double), ("is_put", C.c_int), ("is_american", C.c_int), ("basis", C.c_int),
                ("degree", C.c_int), ("scaling", C.c_int), ("want_regression", C.c_int),
                ("want_exercise_steps", C.c_int), ("want_svd", C.c_int), ("state_f32", C.c_int)]


class LsmSteps(C.Structure):
    _fields_ = [("gamma", c_double_p), ("beta", c_double_p), ("sv", c_double_p), ("mean_x", c_double_p),
                ("std_x", c_double_p), ("rank", c_int_p), ("pivot_loss", c_double_p)]


class LsmTiming(C.Structure):
    _fields_ = [("total_ms", C.c_float), ("step_kernel_ms", C.c_float), ("solve_kernel_ms", C.c_float),
                ("step_launches", C.c_int), ("solve_launches", C.c_int), ("other_launches", C.c_int),
                ("sweep_kind", C.c_int)]


# name -> (restype, argtypes); the CPU-only test tier checks every symbol of include/amc.h is exported
PROTOTYPES = {
    "amc_last_error": (C.c_char_p, []),
    "amc_version": (C.c_int, []),
    "amc_ctx_create": (C.c_int, [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "amc_ctx_destroy": (C.c_int, [C.c_void_p]),
    "amc_ctx_sync": (C.c_int, [C.c_void_p]),
    "amc_ctx_device_info": (C.c_int, [C.c_void_p, c_int_p, c_int_p, c_int_p, c_int64_p]),
    "amc_comm_unique_id": (C.c_int, [C.c_char_p]),
    "amc_comm_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_char_p]),
    "amc_comm_info": (C.c_int, [C.c_void_p, c_int_p, c_int_p]),
    "amc_comm_transport": (C.c_int, [C.c_void_p, c_int_p]),
    "amc_comm_allreduce_host": (C.c_int, [C.c_void_p, c_double_p, C.c_int]),
    "amc_paths_generate": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int64,
                                     C.c_int64, C.c_int64, C.c_int, C.c_uint64, C.POINTER(C.c_void_p)]),
    "amc_paths_generate_lean": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int64,
                                          C.c_int64, C.c_int64, C.c_uint64, C.POINTER(C.c_void_p)]),
    "amc_paths_from_normals": (C.c_int, [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double,
                                         C.c_int, C.c_int64, C.c_int64, C.c_int, C.POINTER(C.c_void_p)]),
    "amc_paths_from_normals_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double,
                                             C.c_int, C.c_int64, C.c_int64, C.c_int, C.POINTER(C.c_void_p)]),
    "amc_paths_from_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int,
                                      C.POINTER(C.c_void_p)]),
    "amc_paths_free": (C.c_int, [C.c_void_p]),
    "amc_paths_info": (C.c_int, [C.c_void_p, c_int64_p, c_int64_p, c_int_p, c_int_p, c_int64_p]),
    "amc_paths_column": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "amc_paths_rows": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
    "amc_paths_column_maps": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "amc_lsm_price": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(LsmSpec), c_double_p, C.POINTER(LsmSteps),
                                C.c_void_p, C.c_void_p, C.POINTER(LsmTiming), C.c_int]),
    "amc_lsm_price_with_hits": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(LsmSpec), C.c_void_p, c_double_p,
                                          C.POINTER(LsmSteps), C.c_void_p, C.c_void_p, C.POINTER(LsmTiming), C.c_int]),
    "amc_paths_gather_steps": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "amc_lsm_price_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(LsmSpec), C.c_int, c_double_p, C.c_void_p,
                                      C.POINTER(LsmTiming), C.c_int]),
    "amc_continuation": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "amc_ccr_exposures": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_void_p,
                                    C.c_void_p, C.c_void_p]),
    "amc_percentiles": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_void_p]),
    "amc_intrinsic_value": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_int, C.c_void_p]),
    "amc_regression_fit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int,
                                     C.c_double, C.c_void_p, C.c_void_p, c_int_p]),
    "amc_estimate_continuation": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                            C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p]),
    "amc_apply_exercise": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_int64, C.c_int64]),
    "amc_basis_matrix": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p]),
    "amc_barrier_hit_matrix": (C.c_int, [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]),
    "amc_selftest_philox": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "amc_selftest_normals": (C.c_int, [C.c_void_p, C.c_int, C.c_uint64, C.c_int64, C.c_int, C.c_int, C.c_double,
                                       C.c_double, C.c_void_p, C.c_void_p]),
}

_lib = None


def lib():
    """Load libamc.so (once).  Raises if it has not been built -- there is no Python/CPU substitute."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `python -m american_monte_carlo_b200.build` "
                              "(needs nvcc); american_monte_carlo_b200 has no CPU fallback")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc):
    """Map a nonzero status to the exception type the reference would raise (ValueError for bad arguments)."""
    if rc == 0:
        return
    msg = lib().amc_last_error().decode("utf-8", "replace")
    if rc == ERR_VALUE:
        raise ValueError(msg)
    raise RuntimeError(f"libamc error {rc}: {msg}")
