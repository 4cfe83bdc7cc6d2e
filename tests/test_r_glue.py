"""The reference-side binding, executed without R.

r/gpedm_b200_glue.c is the `.Call` glue an R maintainer would compile into the package (same
registration pattern as the reference's src/RcppExports.cpp:49-59).  R is not installed in the build
image, so the glue is built against tests/r_stub/ - a ~300-line stand-in for the ~30 R C-API functions it
uses - and every registered entry is invoked BY NAME through the registered table, as R's `.Call` does,
and compared with the direct ctypes path.  What this cannot cover is R's own evaluator (r/overlay.R)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUB_SO = os.path.join(ROOT, "build", "libgpedmb200_rstub.so")

REALSXP, INTSXP, LGLSXP, VECSXP, EXTPTRSXP = 14, 13, 10, 19, 22


@pytest.fixture(scope="module")
def R():
    """Compile glue + fake runtime against libgpedm_b200.so and 'dyn.load' it."""
    from gpedm_b200 import _lib
    os.makedirs(os.path.dirname(STUB_SO), exist_ok=True)
    subprocess.run(["gcc", "-shared", "-fPIC", "-O1", "-Wall", "-Werror", "-Wno-cast-function-type",
                    "-I" + os.path.join(ROOT, "tests", "r_stub"), "-I" + os.path.join(ROOT, "include"),
                    "-o", STUB_SO, os.path.join(ROOT, "r", "gpedm_b200_glue.c"),
                    os.path.join(ROOT, "tests", "r_stub", "r_stub_runtime.c"),
                    "-L" + os.path.dirname(_lib.LIB_PATH), "-lgpedm_b200", "-Wl,-rpath," + os.path.dirname(_lib.LIB_PATH)],
                   check=True)
    L = C.CDLL(STUB_SO)
    vp = C.c_void_p
    for name, res, args in [("rstub_load", C.c_int, []), ("rstub_dynamic_symbols", C.c_int, []),
                            ("rstub_entry_name", C.c_char_p, [C.c_int]), ("rstub_entry_nargs", C.c_int, [C.c_int]),
                            ("rstub_nil", vp, []), ("rstub_real", vp, [vp, C.c_ssize_t, C.c_int, C.c_int]),
                            ("rstub_int", vp, [vp, C.c_ssize_t, C.c_int]), ("rstub_type", C.c_int, [vp]),
                            ("rstub_length", C.c_long, [vp]), ("rstub_nrow", C.c_int, [vp]), ("rstub_ncol", C.c_int, [vp]),
                            ("rstub_data", vp, [vp]), ("rstub_list_get", vp, [vp, C.c_char_p]),
                            ("rstub_list_elt", vp, [vp, C.c_long]), ("rstub_last_error", C.c_char_p, []),
                            ("rstub_protect_depth", C.c_int, []),
                            ("rstub_dot_call", C.c_int, [C.c_char_p, C.c_int, C.POINTER(vp), C.POINTER(vp)]),
                            ("rstub_gc", C.c_int, [])]:
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    return Rstub(L)


class RError(RuntimeError):
    pass


class Rstub:
    """Python face of the fake runtime: R values in, R values out."""

    def __init__(self, L):
        self.L = L
        self.nentries = L.rstub_load()

    def real(self, a):
        a = np.asarray(a, dtype=np.float64)
        f = np.asfortranarray(a)
        nrow, ncol = (a.shape[0], a.shape[1]) if a.ndim == 2 else (0, -1)
        return self.L.rstub_real(f.ctypes.data, f.size, nrow, ncol)

    def int(self, a, logical=False):
        a = np.ascontiguousarray(np.atleast_1d(np.asarray(a, dtype=np.int32)))
        return self.L.rstub_int(a.ctypes.data, a.size, 1 if logical else 0)

    def nil(self):
        return self.L.rstub_nil()

    def call(self, name, *args):
        arr = (C.c_void_p * len(args))(*args)
        out = C.c_void_p()
        rc = self.L.rstub_dot_call(name.encode(), len(args), arr, C.byref(out))
        if rc != 0:
            raise RError(self.L.rstub_last_error().decode())
        return out.value

    def value(self, s):
        """SEXP -> numpy / dict / None."""
        t = self.L.rstub_type(s)
        n = self.L.rstub_length(s)
        if t == 0:
            return None
        if t in (REALSXP, INTSXP, LGLSXP):
            ct = C.c_double if t == REALSXP else C.c_int
            v = np.ctypeslib.as_array(C.cast(self.L.rstub_data(s), C.POINTER(ct)), shape=(n,)).copy() if n else np.empty(0)
            if self.L.rstub_ncol(s) >= 1 and self.L.rstub_nrow(s) * self.L.rstub_ncol(s) == n and self.L.rstub_nrow(s) != n:
                v = v.reshape((self.L.rstub_nrow(s), self.L.rstub_ncol(s)), order="F")
            return v
        if t == VECSXP:
            return [self.L.rstub_list_elt(s, i) for i in range(n)]
        return s

    def get(self, lst, name):
        e = self.L.rstub_list_get(lst, name.encode())
        assert e is not None, name
        return self.value(e)

    def gc(self):
        return self.L.rstub_gc()


ENTRIES = {"gpedmb200_create": 6, "gpedmb200_likegrad": 6, "gpedmb200_fit": 5, "gpedmb200_vector": 2,
           "gpedmb200_matrix": 2, "gpedmb200_predict_new": 4, "gpedmb200_predict_leaveout": 6,
           "gpedmb200_predict_sequential": 2, "gpedmb200_getcov": 10, "gpedmb200_covinv": 2,
           "gpedmb200_create_lagged": 8, "gpedmb200_fit_stats": 1, "gpedmb200_fit_grid": 8}


def test_registration_table_and_error_path(R):
    """`library.dynam` side: R_init_gpedmb200 registers every entry with its argument count and switches
    dynamic symbol lookup off (as src/RcppExports.cpp:55-58 does); `.Call` with a wrong name or argument
    count fails like R's; an error status of the C ABI arrives as an R error (Rf_error), not a crash."""
    got = {R.L.rstub_entry_name(i).decode(): R.L.rstub_entry_nargs(i) for i in range(R.nentries)}
    assert got == ENTRIES
    assert R.L.rstub_dynamic_symbols() == 0
    with pytest.raises(RError, match="not in load table"):
        R.call("gpedmb200_nope", R.nil())
    with pytest.raises(RError, match="Incorrect number of arguments"):
        R.call("gpedmb200_vector", R.nil())
    # a C-ABI error status -> Rf_error with the library's message (no GPU needed: bad arguments are rejected first)
    X = np.zeros((4, 2))
    with pytest.raises(RError, match="libgpedm_b200 status -1"):
        R.call("gpedmb200_create", R.real(X), R.real(np.zeros(4)), R.int([0, 0, 5, 0]), R.int(1), R.nil(), R.int(0))
    with pytest.raises(RError, match="gpedm handle has been released|R_ExternalPtrAddr"):
        R.call("gpedmb200_vector", R.real([1.0]), R.int(0))
    assert R.L.rstub_protect_depth() == 0
    R.gc()


@pytest.mark.gpu
def test_every_dot_call_entry_matches_the_direct_path(R, gpu_lib):
    import gpedm_b200 as G
    from gpedm_b200 import _lib, batch, synth
    Y, X, pop = synth.hierarchical_thetalog(npop=2, n=90, E=2, seed0=11)
    N, d = X.shape
    time = np.concatenate([np.arange(np.sum(pop == k)) for k in range(2)]).astype(float)
    p0 = batch.default_initparst(d, single_pop=False)
    fixed = np.full(d + 2, np.nan)
    m = G.GPModel(Y, X, pop)
    ref = m.fit_rprop(p0)
    # create + fit (fmingrad_Rprop, R/GPfunctions.R:522-584)
    h = R.call("gpedmb200_create", R.real(X), R.real(Y), R.int(m.codes), R.int(2), R.nil(), R.int(0))
    assert R.L.rstub_type(h) == EXTPTRSXP
    fit = R.call("gpedmb200_fit", h, R.real(p0), R.real([1.0]), R.real(fixed), R.real([np.nan]))
    assert int(R.get(fit, "iter")[0]) == ref["iter"]
    for k in ("pars", "parst", "grad"):
        np.testing.assert_array_equal(R.get(fit, k), ref[k])
    for k in ("nllpost", "lp", "SS", "logdet", "lnL_LOO", "df"):
        assert R.get(fit, k)[0] == ref[k]
    # getlikegrad (:586-750), gradient-only and full
    for full in (0, 1):
        lg = R.call("gpedmb200_likegrad", h, R.real(ref["parst"]), R.real([1.0]), R.real(fixed), R.real([np.nan]), R.int(full, True))
        one = m.likegrad(ref["parst"], 1.0, None, None, returngradonly=not full)
        assert R.get(lg, "nllpost")[0] == one["nllpost"]
        np.testing.assert_array_equal(R.get(lg, "grad"), one["grad"])
    # resident vectors / matrices (fit object, :490-519)
    for which in (_lib.VEC_MEAN, _lib.VEC_VAR, _lib.VEC_LOO_MEAN):
        np.testing.assert_array_equal(R.value(R.call("gpedmb200_vector", h, R.int(which))), m.vector(which))
    for which in (_lib.MAT_IKVS, _lib.MAT_CD, _lib.MAT_POPSAME):
        got = R.value(R.call("gpedmb200_matrix", h, R.int(which)))
        assert got.shape == (N, N)
        np.testing.assert_array_equal(got, m.matrix(which))
    # predict.GP newdata (:1012-1036)
    rng = np.random.default_rng(0)
    Xn, pn = rng.standard_normal((7, d)), np.array([0, 1, 1, 0, 0, 1, 0], dtype=np.int32)
    pr = R.call("gpedmb200_predict_new", h, R.real(Xn), R.int(pn), R.int(1, True))
    gm, gv, gg = m.predict_new(Xn, pn, True)
    np.testing.assert_array_equal(R.get(pr, "mean"), gm)
    np.testing.assert_array_equal(R.get(pr, "var"), gv)
    np.testing.assert_array_equal(R.get(pr, "grad"), gg)
    assert R.get(R.call("gpedmb200_predict_new", h, R.real(Xn), R.int(pn), R.int(0, True)), "grad") is None
    # loo / lto (:1056-1104, :1140-1194) with gradient, sequential (:1196-1256)
    prim = np.ones(N, dtype=np.int32)
    for method, fn in ((0, lambda: m.predict_loo(1, time, None, True)), (1, lambda: m.predict_lto(time, 1, None, True))):
        lo = R.call("gpedmb200_predict_leaveout", h, R.int(method), R.int(1), R.real(time), R.int(prim), R.int(1, True))
        em, ev, eg = fn()
        np.testing.assert_array_equal(R.get(lo, "mean"), em)
        np.testing.assert_array_equal(R.get(lo, "var"), ev)
        np.testing.assert_array_equal(R.get(lo, "grad"), eg)
    sq = R.call("gpedmb200_predict_sequential", h, R.real(time))
    sm, ss = m.predict_sequential(time)
    np.testing.assert_array_equal(R.get(sq, "ymean"), sm)
    np.testing.assert_array_equal(R.get(sq, "ysd2"), ss)
    # getcov (:779-816) and getcovinv (:818-830)
    phi = ref["pars"][:d]
    cd = R.value(R.call("gpedmb200_getcov", R.real(phi), R.real([ref["pars"][d + 1]]), R.real([ref["pars"][d + 2]]), R.real(X),
                        R.int(m.codes), R.real(Xn), R.int(pn), R.int(2), R.nil(), R.int(0)))
    np.testing.assert_array_equal(cd, G.getcov(phi, ref["pars"][d + 1], ref["pars"][d + 2], X, Xn, pop, pn)["Cd"])
    S = m.matrix(_lib.MAT_SIGMA)
    ci = R.call("gpedmb200_covinv", R.real(S), R.int(0))
    direct = G.getcovinv(S)
    np.testing.assert_array_equal(R.get(ci, "L").T, direct["L"])
    np.testing.assert_array_equal(R.get(ci, "iKVs"), direct["iKVs"])
    # on-device preparation, fit statistics, E x tau grid (vignettes/choosingE.Rmd:63-86)
    y = synth.ricker(150, 3)
    hl = R.call("gpedmb200_create_lagged", R.real(y), R.nil(), R.int(1), R.int(2), R.int(1), R.int(1), R.nil(), R.int(0))
    ml = G.GPModel.from_series(y, 2, 1, scaling="global")
    pl = batch.default_initparst(2, True)
    fl = R.call("gpedmb200_fit", hl, R.real(pl), R.real([1.0]), R.real(np.full(4, np.nan)), R.real([np.nan]))
    rl = ml.fit_rprop(pl)
    assert int(R.get(fl, "iter")[0]) == rl["iter"] and R.get(fl, "nllpost")[0] == rl["nllpost"]
    st = R.value(R.call("gpedmb200_fit_stats", hl))
    es = ml.fit_stats()
    np.testing.assert_array_equal(st, [es["R2_insample"], es["R2_loo"], es["rmse_insample"], es["rmse_loo"]])
    Es, taus = np.array([1, 2, 1, 2], dtype=np.int32), np.array([1, 1, 2, 2], dtype=np.int32)
    gr = R.value(R.call("gpedmb200_fit_grid", R.real(y), R.nil(), R.int(1), R.int(Es), R.int(taus), R.int(1), R.real([1.0]), R.int(1)))
    direct = {(g["E"], g["tau"]): g for g in G.fit_grid(y, [1, 2], [1, 2], ngpus=1)}
    for cell, E_, t_ in zip(gr, Es, taus):
        f, s = R.L.rstub_list_get(cell, b"fit"), R.get(cell, "stats")
        assert int(R.get(f, "iter")[0]) == direct[(E_, t_)]["iter"]
        assert s[1] == direct[(E_, t_)]["R2_loo"]
    # a failing C-ABI call surfaces as an R error, not a crash: non positive definite matrix
    with pytest.raises(RError, match="status 2"):
        R.call("gpedmb200_covinv", R.real(np.array([[1.0, 2.0], [2.0, 1.0]])), R.int(0))
    # gc(): finalizers destroy the device handles; a released handle is refused
    assert R.L.rstub_protect_depth() == 0
    assert R.gc() == 2
    m.close()
    ml.close()


def test_overlay_edit_script_applies_to_the_reference():
    """r/predict_GP.ed swaps the numeric blocks of fitGP / predict.GP (R/GPfunctions.R:460-461, :884,
    :1012-1030, :1057-1086, :1144-1176, :1203-1245) for the overlay's helpers.  Apply it to the reference
    source (present in the build container only), check the anchors, the result and the brace balance;
    a different revision of the file must be refused."""
    import importlib.util
    ref = "/root/reference/R/GPfunctions.R"
    if not os.path.exists(ref):
        pytest.skip("reference sources are not on this machine")
    spec = importlib.util.spec_from_file_location("apply_overlay", os.path.join(ROOT, "r", "apply_overlay.py"))
    ao = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ao)
    script = open(os.path.join(ROOT, "r", "predict_GP.ed")).read()
    src = open(ref).read()
    out = ao.apply(src, script)
    assert out.count("{") == out.count("}") and out.count("(") == out.count(")")
    body = out[out.index("predict.GP=function"):out.index("#unscale predictions", out.index("predict.GP=function"))]
    for helper in (".predict_new(", ".predict_loo(", ".predict_lto(", ".predict_sequential("):
        assert body.count(helper) == 1, helper
    # no refactorisation loop and no host copy of the N x N matrices is left in predict.GP's numeric part
    body = "\n".join(ln for ln in body.splitlines() if not ln.lstrip().startswith("#"))  # (a commented-out lfo block)
    assert "getcovinv(" not in body and "object$covm" not in body and "getGPgrad(" not in body
    assert "sqrt(output$covdiag)" in out and "sqrt(diag(output$cov))" not in out
    # the helpers the script calls exist in the overlay, and the overlay binds all four covm matrices
    overlay = open(os.path.join(ROOT, "r", "overlay.R")).read()
    for helper in (".predict_new <- function", ".predict_loo <- function", ".predict_lto <- function",
                   ".predict_sequential <- function", "fmingrad_Rprop <- function", "getcov <- function", "getcovinv <- function"):
        assert helper in overlay, helper
    assert all(k in overlay for k in ("Cd = 0L", "Sigma = 1L", "iKVs = 2L", "popsame = 3L", "makeActiveBinding"))
    # every .Call target named in the overlay is a registered entry with that many arguments
    import re
    for name, args in re.findall(r'\.Call\("(gpedmb200_\w+)"((?:[^()]|\((?:[^()]|\([^()]*\))*\))*)\)', overlay):
        depth, n = 0, 0
        for ch in args:
            depth += ch in "([{"
            depth -= ch in ")]}"
            n += (ch == "," and depth == 0)
        assert ENTRIES[name] == n, (name, n)
    with pytest.raises(SystemExit):
        ao.apply(src.replace("iKVs=object$covm$iKVs", "iKVs = object$covm$iKVs"), script)
