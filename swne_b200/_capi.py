"""ctypes binding of include/swne_nmf.h (libswne_nmf.so).  No torch here: plain pointers and sizes.

The library has no CPU fallback; loading works without a GPU (so the ABI can be checked on a CPU box) but
every compute entry point fails with SWNE_ERR_CUDA there.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libswne_nmf.so")

SWNE_OK = 0
SWNE_WARN_NOT_CONVERGED = 1
SWNE_ERR_BAD_ARG = -1
SWNE_ERR_CUDA = -2
SWNE_ERR_NCCL = -3
SWNE_ERR_NO_MATRIX = -4
SWNE_ERR_UNSUPPORTED = -5
SWNE_ERR_INTERRUPTED = -6
LOSS = {"mse": 0, "mkl": 1}
ENGINE = {"auto": 0, "simt": 1, "tc": 2}
ENGINE_NAME = {0: "auto", 1: "simt", 2: "tc"}
VARIANT_KL_L2_LATE = 1      # include/swne_nmf.h SWNE_VARIANT_KL_L2_LATE
TRANSFORM = {"none": 0, "log": 1, "ft": 2}     # ScaleCounts(method = ...), include/swne_nmf.h SWNE_TRANSFORM_*
INIT = {"given": 0, "random": 1, "nnsvd": 2, "ica": 3, "ica_fast": 4}
MAX_K = 64
PROF_SLOTS = 8
PROF_NAMES = ("contract_H_or_pass", "solve_H", "contract_W", "solve_W", "gram_loss", "mkl_update", "allreduce", "other")

# every symbol include/swne_nmf.h declares (tests check the .so exports all of them)
SYMBOLS = (
    "swne_nmf_last_error", "swne_nmf_abi_version", "swne_nmf_device_count", "swne_nmf_default_params",
    "swne_nmf_create", "swne_nmf_destroy", "swne_nmf_set_progress", "swne_nmf_comm_unique_id", "swne_nmf_comm_init", "swne_nmf_comm_info",
    "swne_nmf_set_matrix_dense_f64", "swne_nmf_set_matrix_dense_f32", "swne_nmf_set_matrix_csc_f64",
    "swne_nmf_set_matrix_csc_f32", "swne_nmf_set_matrix_csc_counts", "swne_nmf_set_matrix_dense_f32_device", "swne_nmf_matrix_info", "swne_nmf_fit",
    "swne_nmf_fit_dense_f64", "swne_nmf_fit_csc_f64", "swne_run_nmf", "swne_find_num_factors", "swne_nnlm", "swne_nnsvd_init", "swne_ica_init", "swne_zero_replace",
    "swne_sample_coords", "swne_factor_distance", "swne_recon_error",
)


class SwneError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libswne_nmf status %d: %s" % (code, msg))
        self.code = code


class Params(ctypes.Structure):
    _fields_ = [("k", ctypes.c_int), ("alpha", ctypes.c_double * 3), ("beta", ctypes.c_double * 3),
                ("loss", ctypes.c_int), ("max_iter", ctypes.c_int), ("rel_tol", ctypes.c_double),
                ("inner_max_iter", ctypes.c_int), ("inner_rel_tol", ctypes.c_double), ("trace", ctypes.c_int),
                ("verbose", ctypes.c_int), ("engine", ctypes.c_int), ("full_traces", ctypes.c_int),
                ("profile", ctypes.c_int), ("nnlm_variant", ctypes.c_int)]


class Result(ctypes.Structure):
    _fields_ = [("n_iteration", ctypes.c_int), ("n_trace", ctypes.c_int), ("not_converged", ctypes.c_int),
                ("engine_used", ctypes.c_int), ("loop_ms", ctypes.c_double), ("upload_ms", ctypes.c_double),
                ("download_ms", ctypes.c_double), ("gpu_launches", ctypes.c_longlong),
                ("prof_ms", ctypes.c_double * PROF_SLOTS), ("prof_launches", ctypes.c_longlong * PROF_SLOTS)]


PROGRESS_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                               ctypes.c_double, ctypes.c_double)

_lib = None


def load(build_if_missing=True):
    """dlopen libswne_nmf.so (building it first when the in-tree .so is absent or stale)."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_missing:
        from . import build as _build
        _build.build()
    if not os.path.exists(LIB_PATH):
        raise SwneError(SWNE_ERR_CUDA, "libswne_nmf.so is missing (run python -m swne_b200.build); there is no fallback")
    lib = ctypes.CDLL(LIB_PATH)
    vp, ip, dp, fp = ctypes.c_void_p, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_float)
    c_int, c_dbl = ctypes.c_int, ctypes.c_double
    lib.swne_nmf_last_error.restype = ctypes.c_char_p
    lib.swne_nmf_abi_version.restype = c_int
    lib.swne_nmf_device_count.argtypes = [ip]
    lib.swne_nmf_default_params.argtypes = [ctypes.POINTER(Params), c_int]
    lib.swne_nmf_default_params.restype = None
    lib.swne_nmf_create.argtypes = [ctypes.POINTER(vp), c_int]
    lib.swne_nmf_destroy.argtypes = [vp]
    lib.swne_nmf_set_progress.argtypes = [vp, PROGRESS_FN, vp]
    lib.swne_nmf_comm_unique_id.argtypes = [vp]
    lib.swne_nmf_comm_init.argtypes = [vp, vp, c_int, c_int]
    lib.swne_nmf_comm_info.argtypes = [vp, ip, ip, ip]
    lib.swne_nmf_set_matrix_dense_f64.argtypes = [vp, dp, c_int, c_int]
    lib.swne_nmf_set_matrix_dense_f32.argtypes = [vp, fp, c_int, c_int]
    lib.swne_nmf_set_matrix_csc_f64.argtypes = [vp, ip, ip, dp, c_int, c_int]
    lib.swne_nmf_set_matrix_csc_f32.argtypes = [vp, ip, ip, fp, c_int, c_int]
    lib.swne_nmf_set_matrix_csc_counts.argtypes = [vp, ip, ip, dp, c_int, c_int, dp, dp, c_int]
    lib.swne_nmf_set_matrix_dense_f32_device.argtypes = [vp, vp, c_int, c_int]
    lib.swne_nmf_matrix_info.argtypes = [vp, ip, ip, ctypes.POINTER(ctypes.c_longlong), dp, dp]
    lib.swne_nmf_fit.argtypes = [vp, ctypes.POINTER(Params), dp, dp, dp, dp, dp, dp, c_int, ctypes.POINTER(Result)]
    lib.swne_nmf_fit_dense_f64.argtypes = [vp, ctypes.POINTER(Params), dp, c_int, c_int, dp, dp, dp, dp, dp, dp, c_int,
                                           ctypes.POINTER(Result)]
    lib.swne_nmf_fit_csc_f64.argtypes = [vp, ctypes.POINTER(Params), ip, ip, dp, c_int, c_int, dp, dp, dp, dp, dp, dp,
                                         c_int, ctypes.POINTER(Result)]
    lib.swne_run_nmf.argtypes = [vp, ctypes.POINTER(Params), c_int, ctypes.c_uint64, dp, dp, dp, dp, dp, dp, c_int, ctypes.POINTER(Result)]
    lib.swne_find_num_factors.argtypes = [vp, ip, c_int, c_int, c_int, ctypes.c_uint64, c_int, dp, ip, dp]
    lib.swne_nnlm.argtypes = [vp, c_int, c_int, dp, dp, dp, c_int, c_int, c_dbl, ctypes.POINTER(ctypes.c_longlong)]
    lib.swne_nnsvd_init.argtypes = [vp, c_int, c_int, c_int, ctypes.c_uint64, dp, dp, dp]
    lib.swne_ica_init.argtypes = [vp, c_int, c_int, ctypes.c_uint64, dp, dp]
    lib.swne_zero_replace.argtypes = [dp, ctypes.c_size_t, c_dbl, dp, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]
    lib.swne_sample_coords.argtypes = [vp, dp, c_int, c_int, dp, c_dbl, c_int, dp]
    lib.swne_factor_distance.argtypes = [vp, dp, c_int, c_int, c_int, dp]
    lib.swne_recon_error.argtypes = [vp, dp, dp, c_int, dp, dp]
    for name in SYMBOLS:
        fn = getattr(lib, name)
        if name not in ("swne_nmf_last_error", "swne_nmf_default_params"):
            fn.restype = c_int
    _lib = lib
    return lib


def check(code, allow_warn=True):
    if code == SWNE_OK or (allow_warn and code == SWNE_WARN_NOT_CONVERGED):
        return code
    raise SwneError(code, load().swne_nmf_last_error().decode("utf-8", "replace"))


def dptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def fptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def iptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


def pad3(v):
    """NNLM pads alpha/beta with zeros to [L2, angle, L1]."""
    v = np.atleast_1d(np.asarray(v, dtype=np.float64)).ravel()
    return np.concatenate([v, np.zeros(3)])[:3]
