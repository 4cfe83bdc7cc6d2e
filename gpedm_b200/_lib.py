"""ctypes binding of libgpedm_b200.so (C ABI declared in include/gpedm.h).

The product path has NO CPU fallback: if the shared library is missing or no CUDA device is
present, calls raise `GPEDMError` loudly instead of silently computing on the host.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgpedm_b200.so")
SRC = os.path.join(_HERE, "csrc", "gpedm_core.cu")

GPEDM_MAX_D = 32
GPEDM_MAX_P = GPEDM_MAX_D + 3
NPHASE = 7
PHASE_NAMES = ["assemble", "potrf", "trtri", "lauum", "vec", "trace", "total"]

VEC_MEAN, VEC_VAR, VEC_ALPHA, VEC_DIAG_IKVS, VEC_LOO_MEAN, VEC_LOO_VAR = range(6)
MAT_CD, MAT_SIGMA, MAT_IKVS, MAT_POPSAME, MAT_L, MAT_LINV = range(6)


class GPEDMError(RuntimeError):
    """Raised for any non-zero status of the C ABI (status code in `.status`)."""

    def __init__(self, status, msg):
        super().__init__("libgpedm_b200 status %d: %s" % (status, msg))
        self.status = status


class Result(C.Structure):
    _fields_ = [
        ("nllpost", C.c_double), ("lp", C.c_double), ("like", C.c_double), ("SS", C.c_double),
        ("logdet", C.c_double), ("lnL_LOO", C.c_double), ("df", C.c_double), ("gradnorm", C.c_double),
        ("status", C.c_int32), ("iter", C.c_int32), ("nevals", C.c_int32), ("p", C.c_int32),
        ("pars", C.c_double * GPEDM_MAX_P), ("parst", C.c_double * GPEDM_MAX_P),
        ("grad", C.c_double * GPEDM_MAX_P),
    ]

    def to_dict(self):
        p = self.p
        return dict(nllpost=self.nllpost, lp=self.lp, like=self.like, SS=self.SS, logdet=self.logdet,
                    lnL_LOO=self.lnL_LOO, df=self.df, gradnorm=self.gradnorm, status=self.status,
                    iter=self.iter, nevals=self.nevals,
                    pars=np.array(self.pars[:p]), parst=np.array(self.parst[:p]), grad=np.array(self.grad[:p]))


class RpropOptions(C.Structure):
    _fields_ = [("delta0", C.c_double), ("deltamin", C.c_double), ("deltamax", C.c_double),
                ("eta_minus", C.c_double), ("eta_plus", C.c_double), ("maxiter", C.c_int32),
                ("tol_grad", C.c_double), ("tol_df", C.c_double)]


class FitDesc(C.Structure):
    _fields_ = [("N", C.c_int64), ("d", C.c_int32), ("np", C.c_int32),
                ("X", C.POINTER(C.c_double)), ("Y", C.POINTER(C.c_double)), ("pop", C.POINTER(C.c_int32)),
                ("rhomat", C.POINTER(C.c_double)), ("initparst", C.POINTER(C.c_double)),
                ("fixedpars", C.POINTER(C.c_double)), ("modeprior", C.c_double), ("rhofixed", C.c_double),
                ("want_loo", C.c_int32), ("loo_mean", C.POINTER(C.c_double)), ("loo_var", C.POINTER(C.c_double)),
                ("insamp_mean", C.POINTER(C.c_double)), ("insamp_var", C.POINTER(C.c_double))]


class FitStats(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("n", "obs_mean", "ss_total", "ss_insample", "ss_loo", "R2_insample",
                                          "R2_loo", "rmse_insample", "rmse_loo")]

    def to_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


SCALE_CODES = {"none": 0, "global": 1, "local": 2}

# every symbol include/gpedm.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "gpedm_abi_version", "gpedm_last_error", "gpedm_device_count", "gpedm_default_rprop_options",
    "gpedm_create", "gpedm_set_data", "gpedm_destroy", "gpedm_n", "gpedm_d", "gpedm_likegrad", "gpedm_fit_rprop", "gpedm_get_vector",
    "gpedm_get_matrix", "gpedm_getcov", "gpedm_covinv", "gpedm_predict_new", "gpedm_predict_loo", "gpedm_predict_lto",
    "gpedm_predict_sequential", "gpedm_fit_batch", "gpedm_last_gather_status", "gpedm_create_lagged", "gpedm_lagged_info",
    "gpedm_get_fit_stats", "gpedm_fit_grid", "gpedm_last_phase_ms", "gpedm_launch_count",
    "gpedm_fp64_peak", "gpedm_bench_evals",
]

NVCC_FLAGS = ["-O3", "-std=c++17", "--shared", "-Xcompiler", "-fPIC",
              "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo"]


def build(force=False, verbose=False):
    """Compile libgpedm_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    srcs = [os.path.join(_HERE, "csrc", f) for f in os.listdir(os.path.join(_HERE, "csrc"))]
    srcs.append(os.path.join(_HERE, "..", "include", "gpedm.h"))
    if not force and os.path.exists(LIB_PATH):
        if os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(s) for s in srcs):
            return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB_PATH, SRC, "-ldl"]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return LIB_PATH


_lib = None


def lib():
    """Load the shared library (never builds implicitly on the GPU path; fails loudly)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GPEDMError(-2, "libgpedm_b200.so not found at %s - run `python -c 'import __graft_entry__ as g; "
                             "g.build()'` (no CPU fallback exists)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    L.gpedm_abi_version.restype = C.c_int
    L.gpedm_last_error.restype = C.c_char_p
    L.gpedm_device_count.restype = C.c_int
    L.gpedm_default_rprop_options.argtypes = [C.POINTER(RpropOptions)]
    L.gpedm_create.argtypes = [C.c_int, C.c_int64, C.c_int32, C.c_int32, dp, dp, ip, dp, C.POINTER(C.c_void_p)]
    L.gpedm_set_data.argtypes = [C.c_void_p, dp, dp, ip]
    L.gpedm_destroy.argtypes = [C.c_void_p]
    L.gpedm_n.argtypes = [C.c_void_p]
    L.gpedm_n.restype = C.c_int64
    L.gpedm_d.argtypes = [C.c_void_p]
    L.gpedm_d.restype = C.c_int32
    L.gpedm_likegrad.argtypes = [C.c_void_p, dp, C.c_double, dp, C.c_double, C.c_int, C.POINTER(Result)]
    L.gpedm_fit_rprop.argtypes = [C.c_void_p, dp, C.c_double, dp, C.c_double, C.POINTER(RpropOptions),
                                  C.POINTER(Result)]
    L.gpedm_get_vector.argtypes = [C.c_void_p, C.c_int, dp]
    L.gpedm_get_matrix.argtypes = [C.c_void_p, C.c_int, dp]
    L.gpedm_getcov.argtypes = [C.c_int, C.c_int32, dp, C.c_double, C.c_double, C.c_int64, dp, ip, C.c_int64, dp, ip,
                               C.c_int32, dp, dp]
    L.gpedm_covinv.argtypes = [C.c_int, C.c_int64, dp, dp, dp]
    L.gpedm_predict_new.argtypes = [C.c_void_p, C.c_int64, dp, ip, dp, dp, dp]
    L.gpedm_predict_loo.argtypes = [C.c_void_p, C.c_int32, dp, ip, dp, dp, dp]
    L.gpedm_predict_lto.argtypes = [C.c_void_p, C.c_int32, dp, ip, dp, dp, dp]
    L.gpedm_predict_sequential.argtypes = [C.c_void_p, dp, dp, dp]
    L.gpedm_fit_batch.argtypes = [C.c_int32, C.POINTER(FitDesc), C.c_int32, ip, C.POINTER(RpropOptions),
                                  C.POINTER(Result)]
    i64p = C.POINTER(C.c_int64)
    L.gpedm_create_lagged.argtypes = [C.c_int, C.c_int64, dp, ip, C.c_int32, C.c_int32, C.c_int32, C.c_int32, dp,
                                      C.POINTER(C.c_void_p)]
    L.gpedm_lagged_info.argtypes = [C.c_void_p, i64p, dp, dp, dp, dp, ip]
    L.gpedm_get_fit_stats.argtypes = [C.c_void_p, C.POINTER(FitStats)]
    L.gpedm_fit_grid.argtypes = [C.c_int64, dp, ip, C.c_int32, C.c_int32, ip, ip, C.c_int32, C.c_double, C.c_int32, ip,
                                 C.POINTER(RpropOptions), C.POINTER(Result), C.POINTER(FitStats)]
    L.gpedm_last_phase_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    L.gpedm_launch_count.argtypes = [C.c_void_p]
    L.gpedm_launch_count.restype = C.c_int64
    L.gpedm_fp64_peak.argtypes = [C.c_int, C.c_int, dp]
    L.gpedm_bench_evals.argtypes = [C.c_void_p, dp, C.c_double, C.c_int, C.POINTER(C.c_float)]
    _lib = L
    return L


def check(status):
    if status != 0:
        raise GPEDMError(status, lib().gpedm_last_error().decode("utf-8", "replace"))


def dptr(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def iptr(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int32))


def f64(a, order="F"):
    return np.require(np.asarray(a, dtype=np.float64), requirements=["ALIGNED", "F" if order == "F" else "C"])
