"""ctypes binding of include/gpcsd_b200.h (libgpcsd_b200.so).

The library is the product: if it is missing, fails to load, or a call returns non-zero, this module
raises -- there is no CPU fallback and nothing here imports ``oracle/``.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_long, c_uint, c_ulonglong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgpcsd_b200.so")

KIND_SE = 0
KIND_MATERN = 1


class GpcsdLibraryError(RuntimeError):
    pass


_P = c_void_p  # device pointers and cudaStream_t travel as integers

# name -> (restype, argtypes).  Mirrors include/gpcsd_b200.h one to one; tests/test_abi.py parses the
# header and checks that every declared symbol is exported and bound here.
SIGNATURES = {
    "gpcsd_abi_version": (c_int, []),
    "gpcsd_last_error": (c_char_p, []),
    "gpcsd_num_sms": (c_int, []),
    "gpcsd_dgemm": (c_int, [c_int, c_int, c_int, c_int, _P, c_long, c_long, _P, c_long, c_long, _P, c_long, c_long,
                            c_int, _P]),
    "gpcsd_project_quad_ws_doubles": (c_long, [c_int, c_int, c_int]),
    "gpcsd_project_quad": (c_int, [c_int, c_int, c_int, _P, c_long, _P, c_long, _P, c_long, _P, _P, _P, _P]),
    "gpcsd_project_quad_strided": (c_int, [c_int, c_int, c_int, _P, c_long, _P, c_long, c_long, _P, c_long, _P, _P, _P, _P]),
    "gpcsd_project_quad_batched_ws_doubles": (c_long, [c_int, c_int, c_int, c_int]),
    "gpcsd_project_quad_batched": (c_int, [c_int, c_int, c_int, c_int, _P, c_long, c_long, _P, c_long, c_long, _P, c_long, _P, _P, _P,
                                           c_long, _P]),
    "gpcsd_wsyrk_ws_doubles": (c_long, [c_int, c_int, c_int]),
    "gpcsd_wsyrk": (c_int, [c_int, c_int, c_int, _P, c_long, c_long, _P, _P, c_long, _P, _P]),
    "gpcsd_wsyrk_batched_ws_doubles": (c_long, [c_int, c_int, c_int, c_int, c_int]),
    "gpcsd_wsyrk_batched": (c_int, [c_int, c_int, c_int, c_int, _P, c_long, c_long, c_long, _P, c_long, _P, _P, c_long, c_long, _P, _P]),
    "gpcsd_wsyrk_pair": (c_int, [c_int, c_int, c_int, _P, c_long, c_long, _P, _P, _P, c_long, _P, _P]),
    "gpcsd_eigh_ws_doubles": (c_long, [c_int, c_long]),
    "gpcsd_eigh": (c_int, [c_int, _P, c_long, _P, c_long, _P, _P, c_long, _P, _P]),
    "gpcsd_eigh_batched_ws_bytes": (c_long, [c_int, c_long, c_int]),
    "gpcsd_eigh_batched": (c_int, [c_int, c_int, _P, c_long, _P, _P, c_long, _P, _P]),
    "gpcsd_tridiag": (c_int, [c_int, c_int, _P, c_long, _P, _P, _P, c_long, _P, _P]),
    "gpcsd_backtransform": (c_int, [c_int, c_int, _P, c_long, _P, _P, c_long, _P]),
    "gpcsd_tridiag_eig_ws_doubles": (c_long, [c_int, c_long, c_int]),
    "gpcsd_tridiag_eig": (c_int, [c_int, c_int, _P, _P, _P, _P, c_long, _P, c_long, _P, _P]),
    "gpcsd_eigh_dc_ws_doubles": (c_long, [c_int, c_long, c_int]),
    "gpcsd_eigh_dc": (c_int, [c_int, c_int, _P, c_long, _P, c_long, _P, _P, c_long, _P, _P]),
    "gpcsd_centro_split": (c_int, [c_int, _P, c_long, _P, c_long, _P, c_long, _P]),
    "gpcsd_centro_assemble": (c_int, [c_int, _P, c_long, _P, _P, c_long, _P, _P, c_long, _P, _P]),
    "gpcsd_centro_fold": (c_int, [c_int, c_int, c_long, _P, _P, _P]),
    "gpcsd_centro_unfold": (c_int, [c_int, c_int, c_long, _P, _P, _P]),
    "gpcsd_pairsym_split": (c_int, [c_int, _P, c_long, _P, _P, _P, c_long, _P, c_long, _P]),
    "gpcsd_pairsym_assemble": (c_int, [c_int, _P, _P, _P, c_long, _P, _P, c_long, _P, _P, c_long, _P, _P]),
    "gpcsd_pairsym_fold": (c_int, [c_int, _P, _P, c_long, _P, _P, _P]),
    "gpcsd_eig_D": (c_int, [c_int, c_int, _P, _P, _P, c_int, _P, c_long, _P, _P, _P, _P, _P, _P]),
    "gpcsd_fwd_weights_1d": (c_int, [c_int, _P, c_int, _P, _P, c_double, _P, _P, c_long, _P]),
    "gpcsd_fwd_weights_2d": (c_int, [c_int, _P, c_int, c_int, _P, _P, _P, _P, c_double, c_double, _P, _P, c_long, _P]),
    "gpcsd_se_matrix": (c_int, [c_int, _P, c_int, _P, c_double, c_double, c_int, _P, c_long, _P]),
    "gpcsd_se_grid_to_pts": (c_int, [c_int, c_int, _P, _P, c_int, _P, c_double, c_double, _P, c_long, _P]),
    "gpcsd_kt_build": (c_int, [c_int, _P, c_int, _P, c_int, POINTER(c_int), POINTER(c_double), POINTER(c_double), _P,
                               c_long, _P]),
    "gpcsd_kt_grad_ws_doubles": (c_long, [c_int, c_int]),
    "gpcsd_kt_grad": (c_int, [c_int, _P, c_int, POINTER(c_int), POINTER(c_double), POINTER(c_double), _P, c_long, _P,
                              _P, _P]),
    "gpcsd_grad_core": (c_int, [c_int, _P, c_long, _P, c_long, _P, _P, _P, c_double, c_double, c_double, _P, c_long,
                                _P]),
    "gpcsd_add_diag": (c_int, [c_int, _P, c_long, c_double, _P]),
    "gpcsd_transpose": (c_int, [c_int, c_int, _P, c_long, _P, c_long, _P]),
    "gpcsd_dot_ws_doubles": (c_long, [c_long]),
    "gpcsd_dot": (c_int, [c_int, c_int, _P, c_long, _P, c_long, _P, _P, _P]),
    "gpcsd_sum_arrays": (c_int, [c_long, c_int, POINTER(c_void_p), _P, _P]),
    "gpcsd_sum_vec": (c_int, [c_long, _P, _P, _P]),
    "gpcsd_plan_create": (c_int, [POINTER(c_void_p), c_int, c_int, c_int, POINTER(c_double), POINTER(c_double), c_int,
                                  POINTER(c_double), POINTER(c_double), c_int, POINTER(c_double), POINTER(c_double), c_int,
                                  POINTER(c_int), c_int, c_double, c_double, c_int, c_int, POINTER(c_int), POINTER(c_int), c_int]),
    "gpcsd_plan_destroy": (c_int, [_P]),
    "gpcsd_plan_num_params": (c_int, [_P]),
    "gpcsd_plan_ws_bytes": (c_long, [_P, c_long, c_int]),
    "gpcsd_plan_set_lfp": (c_int, [_P, _P, c_long, c_int, c_double, c_double, _P, c_long]),
    "gpcsd_plan_touch_lfp": (c_int, [_P]),
    "gpcsd_plan_set_graph": (c_int, [_P, c_int]),
    "gpcsd_plan_loglik_grad": (c_int, [_P, c_int, POINTER(c_double), c_int, POINTER(c_double), _P]),
    "gpcsd_plan_enqueue": (c_int, [_P, c_int, POINTER(c_double), c_int, _P]),
    "gpcsd_plan_finish": (c_int, [_P, c_int, POINTER(c_double), _P]),
    "gpcsd_plan_device_result": (c_void_p, [_P]),
    "gpcsd_plan_device_theta": (c_void_p, [_P]),
    "gpcsd_plan_last_launches": (c_long, [_P]),
    "gpcsd_plan_mailbox_bytes": (c_long, [_P, c_int]),
    "gpcsd_plan_set_mailbox": (c_int, [_P, _P, c_long, c_int, c_int]),
    "gpcsd_plan_loglik_grad_factors": (c_int, [_P, POINTER(c_double), _P, _P, _P, _P, c_int, POINTER(c_double), _P]),
    "gpcsd_fwd_operator_1d": (c_int, [c_int, _P, c_int, _P, c_double, c_double, _P, c_long, _P]),
    "gpcsd_fwd_operator_2d": (c_int, [c_int, _P, c_int, _P, c_int, _P, c_double, c_double, _P, c_long, _P]),
    "gpcsd_cholesky": (c_int, [c_int, _P, c_long, _P, _P]),
    "gpcsd_randn": (c_int, [c_long, c_long, c_long, c_ulonglong, c_uint, c_double, c_int, _P, _P]),
    "gpcsd_philox_raw": (c_int, [c_long, c_ulonglong, c_uint, _P, _P]),
    "gpcsd_shift_residual": (c_int, [c_int, c_int, c_int, c_long, _P, c_int, _P, _P, c_int, _P, _P, _P]),
    "gpcsd_per_trial_ws_doubles": (c_long, [c_int, c_int, c_int, c_long, c_int]),
    "gpcsd_quad_per_trial": (c_int, [c_int, c_int, c_int, c_long, _P, _P, c_long, _P, _P, _P]),
    "gpcsd_shift_grad": (c_int, [c_int, c_int, c_int, c_long, _P, c_int, _P, _P, c_int, _P, _P, _P, _P]),
}

_STATUS_CALLS = {n for n, (r, _) in SIGNATURES.items()
                 if r is c_int and n not in ("gpcsd_abi_version", "gpcsd_num_sms", "gpcsd_plan_num_params")}

_lib = None


def load():
    """Load libgpcsd_b200.so (built in-tree by gpcsd_b200/build.py); raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GpcsdLibraryError(
            "libgpcsd_b200.so not found at %s -- run `python -m gpcsd_b200.build` (or __graft_entry__.build()); "
            "there is no CPU fallback" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.gpcsd_abi_version() != 1:
        raise GpcsdLibraryError("ABI version mismatch")
    _lib = lib
    return lib


def call(name, *args):
    """Invoke a status-returning entry point; raise with gpcsd_last_error() on failure."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if name in _STATUS_CALLS and rc != 0:
        raise GpcsdLibraryError("%s failed (%d): %s" % (name, rc, lib.gpcsd_last_error().decode()))
    return rc


def query(name, *args):
    lib = load()
    v = getattr(lib, name)(*args)
    if v < 0:
        raise GpcsdLibraryError("%s failed: %s" % (name, lib.gpcsd_last_error().decode()))
    return v


def c_int_array(vals):
    return (c_int * len(vals))(*[int(v) for v in vals])


def c_double_array(vals):
    return (c_double * len(vals))(*[float(v) for v in vals])
