"""CPU tests of the drop-in boundary: the C-ABI library builds for sm_100a, loads, exports every
symbol include/gpedm.h declares, and refuses loudly (no CPU fallback) when no GPU is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import gpedm_b200
from gpedm_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    _lib.build()
    return gpedm_b200.lib()


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "gpedm.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(gpedm_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_exported(L):
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), "symbol %s declared in include/gpedm.h is not exported" % s
    assert sorted(_lib.EXPORTS) == syms


def test_abi_version_and_struct_layout(L):
    assert L.gpedm_abi_version() == 3
    # gpedm_result: 8 doubles + 4 int32 + 3 * GPEDM_MAX_P doubles
    assert C.sizeof(_lib.Result) == 8 * 8 + 4 * 4 + 3 * 35 * 8
    opt = _lib.RpropOptions()
    L.gpedm_default_rprop_options(C.byref(opt))
    # constants of fmingrad_Rprop, reference R/GPfunctions.R:531-536, :556
    assert (opt.delta0, opt.deltamin, opt.deltamax, opt.eta_minus, opt.eta_plus, opt.maxiter) == \
        (0.1, 1e-6, 50.0, 0.5, 1.2, 200)
    assert (opt.tol_grad, opt.tol_df) == (1e-4, 1e-7)


def test_no_cpu_fallback_without_gpu(L):
    if L.gpedm_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(gpedm_b200.GPEDMError) as e:
        gpedm_b200.GPModel(np.arange(4.0), np.arange(8.0).reshape(4, 2))
    assert e.value.status == -2 and "no CPU fallback" in str(e.value)
    with pytest.raises(gpedm_b200.GPEDMError):
        gpedm_b200.getcov([0.1], 1.0, 0.5, [[0.0], [1.0]], [[0.5]], [1, 1], [1])
    with pytest.raises(gpedm_b200.GPEDMError):
        gpedm_b200.fitGP(y=np.sin(np.arange(30.0)), E=2, tau=1)


def test_argument_validation(L):
    h = C.c_void_p()
    X = np.zeros((4, 2), order="F")
    Y = np.zeros(4)
    pop = np.zeros(4, dtype=np.int32)
    dp, ip = _lib.dptr, _lib.iptr
    assert L.gpedm_create(0, 0, 2, 1, dp(X), dp(Y), ip(pop), None, C.byref(h)) == -1      # N < 1
    assert L.gpedm_create(0, 4, 0, 1, dp(X), dp(Y), ip(pop), None, C.byref(h)) == -1      # d < 1
    assert L.gpedm_create(0, 4, 33, 1, dp(X), dp(Y), ip(pop), None, C.byref(h)) == -1     # d > GPEDM_MAX_D
    assert L.gpedm_create(0, 4, 2, 1, None, dp(Y), ip(pop), None, C.byref(h)) == -1       # NULL X
    bad = np.array([0, 0, 3, 0], dtype=np.int32)
    assert L.gpedm_create(0, 4, 2, 2, dp(X), dp(Y), ip(bad), None, C.byref(h)) == -1      # pop code out of range
    assert b"pop[2]" in L.gpedm_last_error()
    pop2 = np.array([0, 1, 0, 1], dtype=np.int32)
    Rbad = np.array([[1.0, 0.3], [0.3, 0.9]], order="F")                                     # rhomatrix diagonal != 1
    assert L.gpedm_create(0, 4, 2, 2, dp(X), dp(Y), ip(pop2), dp(Rbad), C.byref(h)) == -1
    assert b"unit diagonal" in L.gpedm_last_error()
    assert L.gpedm_likegrad(None, None, 1.0, None, float("nan"), 0, None) == -1
    assert L.gpedm_fit_batch(-1, None, 1, None, None, None) == -1
    assert L.gpedm_create_lagged(0, 10, dp(np.zeros(10)), None, 1, 2, 1, 1, None, None) == -1   # NULL out handle
    assert L.gpedm_fit_grid(10, None, None, 1, 2, None, None, 1, 1.0, 1, None, None, None, None) == -1
    assert L.gpedm_get_fit_stats(None, None) == -1
    assert C.sizeof(_lib.FitStats) == 9 * 8


def test_sources_build_for_sm100a():
    flags = " ".join(_lib.NVCC_FLAGS)
    assert "arch=compute_100a,code=sm_100a" in flags and "-lineinfo" in flags
    assert os.path.exists(_lib.LIB_PATH)


def test_r_glue_type_checks_against_header():
    """R is not installed here, so r/gpedm_b200_glue.c cannot be built; type-check it with gcc against
    include/gpedm.h and minimal stub declarations of the R C API (tests/r_stub/): every C-ABI call in the
    glue must match the header's signature, and every .Call entry must be registered."""
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    glue = os.path.join(root, "r", "gpedm_b200_glue.c")
    r = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Werror", "-Wno-cast-function-type",
                        "-I" + os.path.join(root, "tests", "r_stub"), "-I" + os.path.join(root, "include"), glue],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    src = open(glue).read()
    defined = set(re.findall(r"^SEXP (gpedmb200_\w+)\(", src, flags=re.M))
    registered = set(re.findall(r'\{"(gpedmb200_\w+)", \(DL_FUNC\)&\1, \d+\}', src))
    assert defined == registered and len(defined) >= 13
    for name, nargs in re.findall(r'\{"(gpedmb200_\w+)", \(DL_FUNC\)&\w+, (\d+)\}', src):
        sig = re.search(r"^SEXP %s\(([^)]*)\)" % name, src, flags=re.M | re.S).group(1)
        assert sig.count("SEXP") == int(nargs), name
