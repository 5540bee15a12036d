"""ctypes binding of the C-ABI engine ``libpgb.so`` (``include/pgb.h``).

The library is built in-tree by ``pygam_b200/csrc/build.sh`` (or ``__graft_entry__.build()``).
There is no CPU fallback: if the shared library is missing, importing this module raises.
"""

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libpgb.so")

MAX_MARG = 4
MAX_ORDER = 7

TERM_INTERCEPT, TERM_LINEAR, TERM_SPLINE, TERM_FACTOR, TERM_TENSOR = range(5)
DISTS = {"normal": 0, "binomial": 1, "poisson": 2, "gamma": 3, "inv_gauss": 4}
LINKS = {"identity": 0, "logit": 1, "log": 2, "inverse": 3, "inv_squared": 4}
OK, EINVAL, ECUDA, ENOMEM, ENOTPD, EDIVERGED, EUNSUPPORTED = 0, -1, -2, -3, -4, -5, -6

GS_NUSED, GS_DEV, GS_NCORRECT, GS_ZWZ, GS_NROWS, GS_NONFINITE, GS_COUNT = 0, 1, 2, 3, 4, 5, 8
(SS_N, SS_WDEV, SS_PEARSON, SS_DEV, SS_NCORRECT, SS_SUMY, SS_SUMW, SS_LOGLIK, SS_NULL_WDEV,
 SS_NULL_LOGLIK, SS_NONFINITE) = range(11)
SS_COUNT = 16
MODE_IRLS, MODE_INIT = 0, 1
MODE_PLANNED = 0x100  # or-ed into mode: the workspace holds what pgb_gram_prepare left for this X
SOLVE_OUT_EXTRA = 8
SOLVE_EDOF, SOLVE_COV, SOLVE_PREPARED = 1, 2, 4


class PgbTerm(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("n_marg", C.c_int32),
        ("feature", C.c_int32 * MAX_MARG),
        ("n_splines", C.c_int32 * MAX_MARG),
        ("order", C.c_int32 * MAX_MARG),
        ("periodic", C.c_int32 * MAX_MARG),
        ("edge_lo", C.c_double * MAX_MARG),
        ("edge_hi", C.c_double * MAX_MARG),
        ("by", C.c_int32),
        ("drop_first", C.c_int32),
        ("coef_offset", C.c_int32),
        ("n_coefs", C.c_int32),
    ]


class PgbDebugOpts(C.Structure):
    """test-only plan constraints of pgb_debug_model_create (include/pgb.h)"""
    _fields_ = [("force_general_path", C.c_int32), ("cells_tile_rows", C.c_int32), ("cells_band_tiles", C.c_int32),
                ("cells_threads", C.c_int32), ("cells_kernel", C.c_int32)]


class PgbModelInfo(C.Structure):
    _fields_ = [
        ("m", C.c_int32),
        ("k", C.c_int32),
        ("n_features", C.c_int32),
        ("n_roles", C.c_int32),
        ("tile_rows", C.c_int32),
        ("gram_ctas", C.c_int32),
        ("gram_ws_bytes", C.c_int64),
        ("pass_ws_bytes", C.c_int64),
        ("cell_roles", C.c_int32),
        ("cell_tile_rows", C.c_int32),
        ("cell_threads", C.c_int32),
        ("cell_ring_groups", C.c_int32),
        ("cell_generic_roles", C.c_int32),
        ("reserved0", C.c_int32),
    ]


# every symbol include/pgb.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "pgb_last_error": (C.c_char_p, []),
    "pgb_version": (C.c_int, []),
    "pgb_launch_count": (C.c_longlong, []),
    "pgb_gram_timing": (C.c_int, [_P, C.c_int32]),
    "pgb_gram_last_ms": (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "pgb_model_create": (C.c_int, [C.POINTER(PgbTerm), C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                   C.c_double, C.c_double, C.c_int32, C.POINTER(_P)]),
    "pgb_debug_model_create": (C.c_int, [C.POINTER(PgbTerm), C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                         C.c_double, C.c_double, C.c_int32, C.POINTER(PgbDebugOpts), C.POINTER(_P)]),
    "pgb_model_destroy": (None, [_P]),
    "pgb_model_plan_host": (C.c_int, [C.POINTER(PgbTerm), C.c_int32, C.c_int32, C.POINTER(PgbModelInfo),
                                      C.POINTER(C.c_int32)]),
    "pgb_model_get_info": (C.c_int, [_P, C.POINTER(PgbModelInfo)]),
    "pgb_col_stats_ws_bytes": (C.c_int64, [C.c_int32]),
    "pgb_col_stats": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int32, _P, _P, _P, C.c_int64, _P]),
    "pgb_col_levels": (C.c_int, [_P, C.c_int64, C.c_double, C.c_int32, _P, _P]),
    "pgb_model_matrix": (C.c_int, [_P, _P, C.c_int64, C.c_int64, _P, _P]),
    "pgb_gram_ws_bytes": (C.c_int64, [_P, C.c_int64]),
    "pgb_gram_ws_bytes_streaming": (C.c_int64, [_P, C.c_int64]),
    "pgb_gram_pass": (C.c_int, [_P, C.c_int32, _P, C.c_int64, C.c_int64, _P, _P, _P, _P, _P, C.c_int64, _P]),
    "pgb_gram_prepare": (C.c_int, [_P, _P, C.c_int64, C.c_int64, _P, C.c_int64, _P]),
    "pgb_gram_refine": (C.c_int, [_P, C.c_int64, _P, C.c_int64, _P]),
    "pgb_stats_ws_bytes": (C.c_int64, [_P]),
    "pgb_stats_pass": (C.c_int, [_P, _P, C.c_int64, C.c_int64, _P, _P, _P, C.c_double, C.c_double,
                                 C.c_int32, _P, _P, _P, _P, C.c_int64, _P]),
    "pgb_solve_ws_bytes": (C.c_int64, [C.c_int32, C.c_int32]),
    "pgb_interval_pass": (C.c_int, [_P, _P, C.c_int64, C.c_int64, _P, _P, C.c_int32, _P, _P, _P]),
    "pgb_solve_prepare": (C.c_int, [C.c_int32, C.c_int32, _P, _P, _P, _P, C.c_int32, _P, _P, C.c_int64, _P]),
    "pgb_solve": (C.c_int, [C.c_int32, C.c_int32, _P, _P, C.c_int32, _P, _P, _P, _P, C.c_int32, _P, C.c_int32,
                            C.c_int32, _P, _P, _P, C.c_int64, _P]),
    "pgb_solve_truncated_ws_bytes": (C.c_int64, [C.c_int32]),
    "pgb_solve_truncated": (C.c_int, [C.c_int32, C.c_int32, _P, _P, _P, _P, _P, C.c_int32, _P, _P, _P, C.c_int64, _P]),
    "pgb_chol_ws_bytes": (C.c_int64, [C.c_int32]),
    "pgb_chol_check": (C.c_int, [C.c_int32, _P, _P, _P, C.c_int64, _P]),
    "pgb_chol_check_batch": (C.c_int, [C.c_int32, C.c_int32, _P, _P, _P, C.c_int64, _P]),
    "pgb_gram_pass_host": (C.c_int, [_P, C.c_int32, _P, C.c_int64, C.c_int64, _P, _P, _P, _P, C.c_int64]),
    "pgb_upload_rows": (C.c_int, [_P, C.c_int64, _P, C.c_int64, C.c_int32, C.c_int32]),
}

_lib = None


def load():
    """Load libpgb.so; raises ImportError when it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA engine first "
            "(python -c 'import __graft_entry__ as g; g.build()' or pygam_b200/csrc/build.sh). "
            "pygam_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class PgbError(RuntimeError):
    pass


def check(rc, what=""):
    if rc != OK:
        msg = load().pgb_last_error().decode("utf-8", "replace")
        raise PgbError(f"{what} failed with code {rc}: {msg}")
