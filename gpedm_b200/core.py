"""Host-side mirror of the R-internal numerical functions of GPEDM, backed by libgpedm_b200.

Names and argument meaning follow the reference (R/GPfunctions.R): `getcov` (:779-816),
`getlikegrad` (:586-750), `fmingrad_Rprop` (:522-584), `logit` (:1343-1345).  All N^2 / N^3
work happens on the GPU through the C ABI; this module only marshals arrays.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import GPEDMError, Result, RpropOptions, check, dptr, f64, iptr, lib

VEMIN = SIGMA2MIN = 0.0001  # R/GPfunctions.R:422, 591
VEMAX = SIGMA2MAX = 4.9999
RHOMIN, RHOMAX = 0.0001, 1.0


def logit(x, lo=0.0, hi=1.0):
    """R/GPfunctions.R:1343-1345."""
    x = np.asarray(x, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.log((x - lo) / (hi - x))


def unique_stable(v):
    """R's unique(): values in order of first appearance."""
    seen, out = set(), []
    for e in v:
        if e not in seen:
            seen.add(e)
            out.append(e)
    return out


def pop_codes(pop, levels=None, unseen_ok=False):
    """Population labels -> int32 codes 0..np-1 in order of first appearance (or given levels).
    unseen_ok: labels absent from `levels` (new-data populations the model was not trained on) get the
    sentinel code np, which getcov treats as the reference does (R/GPfunctions.R:796-811): popsame = 0
    against every training row, i.e. the rho cross-covariance, or 1 under a rhomatrix."""
    pop = np.asarray(pop, dtype=object)
    if levels is None:
        levels = unique_stable(pop)
    lut = {u: i for i, u in enumerate(levels)}
    if unseen_ok:
        return np.array([lut.get(p, len(levels)) for p in pop], dtype=np.int32), list(levels)
    try:
        codes = np.array([lut[p] for p in pop], dtype=np.int32)
    except KeyError as e:
        raise ValueError("population %r not present in the training data" % (e.args[0],))
    return codes, list(levels)


def _rhomat_in_code_order(rhomatrix, rhomatrix_names, levels):
    """Reorder a named rhomatrix to code order (reference addresses it by dimnames, :799-808)."""
    if rhomatrix is None:
        return None
    R = np.asarray(rhomatrix, dtype=np.float64)
    names = [str(n) for n in (rhomatrix_names if rhomatrix_names is not None else levels)]
    idx = []
    for u in levels:
        if str(u) not in names:
            raise ValueError("Population names are not all present in rhomatrix.")
        idx.append(names.index(str(u)))
    return f64(R[np.ix_(idx, idx)])


class GPModel:
    """Device-resident state of one GP fit (training set + factorisation products).

    Wraps a `gpedm_handle`; the R overlay would keep the same handle behind an external pointer
    (INTEGRATION.md).  Y (N), X (N x d), pop: labels of any hashable type.
    """

    def __init__(self, Y, X, pop=None, rhomatrix=None, rhomatrix_names=None, device=0):
        Y = f64(Y).ravel()
        X = np.asarray(X, dtype=np.float64)
        if X.ndim == 1:
            X = X[:, None]
        X = f64(X)
        if X.shape[0] != Y.shape[0]:
            raise ValueError("X and Y have different numbers of rows")
        self.N, self.d = X.shape
        if pop is None:
            pop = np.ones(self.N, dtype=np.int64)
        self.codes, self.levels = pop_codes(pop)
        self.np_ = len(self.levels)
        self.rhomat = _rhomat_in_code_order(rhomatrix, rhomatrix_names, self.levels)
        self.Y, self.X = Y, X
        self.device = device
        self._h = C.c_void_p()
        check(lib().gpedm_create(device, self.N, self.d, self.np_, dptr(X), dptr(Y), iptr(self.codes),
                                 dptr(self.rhomat), C.byref(self._h)))

    @classmethod
    def from_series(cls, y, E, tau=1, pop=None, scaling="global", rhomatrix=None, rhomatrix_names=None, device=0):
        """Training set of fitGP(y, E, tau, scaling) prepared ON THE DEVICE from the raw series
        (lags within population, centring/scaling, incomplete-row removal; reference
        R/GPfunctions.R:1824-1857, :359-405, :445-451) - `gpedm_create_lagged`."""
        from ._lib import SCALE_CODES
        y = f64(y).ravel()
        T = len(y)
        self = cls.__new__(cls)
        codes, levels = pop_codes(np.ones(T, dtype=np.int64) if pop is None else pop)
        self.levels, self.np_ = levels, len(levels)
        self.rhomat = _rhomat_in_code_order(rhomatrix, rhomatrix_names, levels)
        self.device = device
        self._h = C.c_void_p()
        check(lib().gpedm_create_lagged(device, T, dptr(y), iptr(codes), self.np_, int(E), int(tau),
                                        SCALE_CODES[scaling], dptr(self.rhomat), C.byref(self._h)))
        self.N, self.d = int(lib().gpedm_n(self._h)), int(E)
        G = self.np_ if scaling == "local" else 1
        rows = np.empty(self.N, dtype=np.int64)
        self.center = np.empty((self.d + 1, G), order="F")
        self.scale = np.empty((self.d + 1, G), order="F")
        self.Y = np.empty(self.N)
        self.X = np.empty((self.N, self.d), order="F")
        self.codes = np.empty(self.N, dtype=np.int32)
        check(lib().gpedm_lagged_info(self._h, rows.ctypes.data_as(C.POINTER(C.c_int64)), dptr(self.center),
                                      dptr(self.scale), dptr(self.Y), dptr(self.X), iptr(self.codes)))
        self.rows = rows
        return self

    def fit_stats(self):
        """In-sample / leave-one-out R2 and rmse in the original units of y, reduced on the device
        (getR2, reference R/GPfunctions.R:1366-1370)."""
        from ._lib import FitStats
        st = FitStats()
        check(lib().gpedm_get_fit_stats(self._h, C.byref(st)))
        return st.to_dict()

    def set_data(self, Y, X, codes=None):
        """Re-upload host Y / X (/ population codes) into the resident handle (same shapes)."""
        Y = f64(Y).ravel()
        X = f64(np.asarray(X, dtype=np.float64).reshape(self.N, self.d))
        codes = self.codes if codes is None else np.ascontiguousarray(codes, dtype=np.int32)
        check(lib().gpedm_set_data(self._h, dptr(X), dptr(Y), iptr(codes)))
        self.Y, self.X, self.codes = Y, X, codes

    # -- lifetime
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().gpedm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- getlikegrad / fmingrad_Rprop
    def _fixed(self, fixedpars):
        if fixedpars is None:
            return None
        fp = f64(fixedpars).ravel()
        if fp.shape[0] != self.d + 2:
            raise ValueError("Length of fixedpars must be number of predictors + 2 (phi, ve, sigma2)")
        return fp

    def likegrad(self, parst, modeprior=1.0, fixedpars=None, rhofixed=None, returngradonly=True):
        parst = f64(parst).ravel()
        if parst.shape[0] != self.d + 3:
            raise ValueError("parst must have d+3 entries")
        fp = self._fixed(fixedpars)
        res = Result()
        check(lib().gpedm_likegrad(self._h, dptr(parst), float(modeprior), dptr(fp),
                                   float("nan") if rhofixed is None else float(rhofixed),
                                   0 if returngradonly else 1, C.byref(res)))
        return res.to_dict()

    def fit_rprop(self, initparst, modeprior=1.0, fixedpars=None, rhofixed=None, options=None):
        initparst = f64(initparst).ravel()
        fp = self._fixed(fixedpars)
        res = Result()
        opt = None
        if options:
            opt = RpropOptions()
            lib().gpedm_default_rprop_options(C.byref(opt))
            for k, v in options.items():
                setattr(opt, k, v)
        check(lib().gpedm_fit_rprop(self._h, dptr(initparst), float(modeprior), dptr(fp),
                                    float("nan") if rhofixed is None else float(rhofixed),
                                    C.byref(opt) if opt is not None else None, C.byref(res)))
        return res.to_dict()

    # -- products of the last full evaluation
    def vector(self, which):
        out = np.empty(self.N)
        check(lib().gpedm_get_vector(self._h, which, dptr(out)))
        return out

    def matrix(self, which):
        out = np.empty((self.N, self.N), order="F")
        check(lib().gpedm_get_matrix(self._h, which, dptr(out)))
        return out

    # -- prediction cores
    def predict_new(self, Xnew, popnew=None, returnGPgrad=False):
        Xnew = np.asarray(Xnew, dtype=np.float64)
        if Xnew.ndim == 1:
            Xnew = Xnew[:, None]
        Xnew = f64(Xnew)
        M = Xnew.shape[0]
        if Xnew.shape[1] != self.d:
            raise ValueError("Xnew must have %d columns" % self.d)
        codes = np.zeros(M, dtype=np.int32) if popnew is None else pop_codes(popnew, self.levels, unseen_ok=True)[0]
        mean, var = np.empty(M), np.empty(M)
        grad = np.empty((M, self.d), order="F") if returnGPgrad else None
        check(lib().gpedm_predict_new(self._h, M, dptr(Xnew), iptr(codes), dptr(mean), dptr(var), dptr(grad)))
        return mean, var, grad

    def _primary(self, primary):
        if primary is None:
            return None, self.N
        pr = np.ascontiguousarray(np.asarray(primary).astype(np.int32))
        return pr, int(pr.sum())

    def predict_loo(self, exclradius=0, time=None, primary=None, returnGPgrad=False):
        pr, nd = self._primary(primary)
        t = None if time is None else f64(time).ravel()
        mean, var = np.empty(nd), np.empty(nd)
        grad = np.empty((nd, self.d), order="F") if returnGPgrad else None
        check(lib().gpedm_predict_loo(self._h, int(exclradius), dptr(t), iptr(pr), dptr(mean), dptr(var), dptr(grad)))
        return mean, var, grad

    def predict_lto(self, time, exclradius=0, primary=None, returnGPgrad=False):
        pr, nd = self._primary(primary)
        t = f64(time).ravel()
        mean, var = np.empty(nd), np.empty(nd)
        grad = np.empty((nd, self.d), order="F") if returnGPgrad else None
        check(lib().gpedm_predict_lto(self._h, int(exclradius), dptr(t), iptr(pr), dptr(mean), dptr(var), dptr(grad)))
        return mean, var, grad

    def predict_sequential(self, time):
        t = f64(time).ravel()
        ymean, ysd2 = np.empty(self.N), np.empty(self.N)
        check(lib().gpedm_predict_sequential(self._h, dptr(t), dptr(ymean), dptr(ysd2)))
        return ymean, ysd2

    # -- measurement
    def phase_ms(self):
        ms = (C.c_float * _lib.NPHASE)()
        check(lib().gpedm_last_phase_ms(self._h, ms))
        return dict(zip(_lib.PHASE_NAMES, [float(x) for x in ms]))

    def launch_count(self):
        return int(lib().gpedm_launch_count(self._h))

    def bench_evals(self, parst, modeprior=1.0, steps=1):
        parst = f64(parst).ravel()
        ms = C.c_float()
        check(lib().gpedm_bench_evals(self._h, dptr(parst), float(modeprior), int(steps), C.byref(ms)))
        return float(ms.value)


def getcov(phi, sigma2, rho, X, X2, pop, pop2, rhomat=None, rhomat_names=None, device=0):
    """R/GPfunctions.R:779-816.  Returns dict(popsame=..., Cd=...) with Cd of shape (nrow(X2), nrow(X))."""
    X = np.asarray(X, dtype=np.float64)
    X2 = np.asarray(X2, dtype=np.float64)
    if X.ndim == 1:
        X = X[:, None]
    if X2.ndim == 1:
        X2 = X2[:, None]
    X, X2 = f64(X), f64(X2)
    phi = f64(phi).ravel()
    codes, levels = pop_codes(pop)
    codes2, _ = pop_codes(pop2, levels, unseen_ok=True)
    R = _rhomat_in_code_order(rhomat, rhomat_names, levels)
    N, M = X.shape[0], X2.shape[0]
    Cd = np.empty((M, N), order="F")
    check(lib().gpedm_getcov(device, X.shape[1], dptr(phi), float(sigma2), float(rho), N, dptr(X), iptr(codes), M,
                             dptr(X2), iptr(codes2), len(levels), dptr(R), dptr(Cd)))
    popsame = (codes2[:, None] == codes[None, :]) * 1.0
    return dict(popsame=popsame, Cd=Cd)


def getcovinv(Sigma, device=0):
    """R/GPfunctions.R:818-830: Cholesky factor and inverse of a covariance matrix.  Returns
    dict(L=<upper factor, like Matrix::chol>, iKVs=<inverse>)."""
    S = f64(Sigma)
    if S.ndim != 2 or S.shape[0] != S.shape[1]:
        raise ValueError("Sigma must be a square matrix")
    n = S.shape[0]
    L = np.empty((n, n), order="F")
    inv = np.empty((n, n), order="F")
    check(lib().gpedm_covinv(device, n, dptr(S), dptr(L), dptr(inv)))
    return dict(L=np.ascontiguousarray(L.T), iKVs=inv)


def getlikegrad(parst, Y, X, pop, modeprior, fixedpars, rhofixed, rhomatrix, D=None, returngradonly=True,
                rhomatrix_names=None, device=0):
    """R/GPfunctions.R:586-750 as a one-shot call (the per-coordinate distance list `D` of the
    reference is accepted for signature compatibility and ignored: distances are recomputed on
    the fly inside the kernels)."""
    m = GPModel(Y, X, pop, rhomatrix, rhomatrix_names, device)
    try:
        out = m.likegrad(parst, modeprior, fixedpars, rhofixed, returngradonly)
        if not returngradonly:
            out["mean"] = m.vector(_lib.VEC_MEAN)
            out["covdiag"] = m.vector(_lib.VEC_VAR)
    finally:
        m.close()
    return out


def fmingrad_Rprop(initparst, Y, X, pop, modeprior, fixedpars, rhofixed, rhomatrix, rhomatrix_names=None,
                   device=0, keep_model=False):
    """R/GPfunctions.R:522-584.  Returns the final full evaluation + `iter`; with keep_model=True the
    device-resident `GPModel` is returned under key "model" for prediction."""
    m = GPModel(Y, X, pop, rhomatrix, rhomatrix_names, device)
    try:
        out = m.fit_rprop(initparst, modeprior, fixedpars, rhofixed)
        out["mean"] = m.vector(_lib.VEC_MEAN)
        out["covdiag"] = m.vector(_lib.VEC_VAR)
    except Exception:
        m.close()
        raise
    if keep_model:
        out["model"] = m
    else:
        m.close()
    return out


def fp64_peak(which=0, device=0):
    """Measured FP64 issue peak in TFLOP/s: which=0 DMMA (tensor pipe), 1 DFMA."""
    v = C.c_double()
    check(lib().gpedm_fp64_peak(device, which, C.byref(v)))
    return float(v.value)


__all__ = ["GPModel", "GPEDMError", "getcov", "getcovinv", "getlikegrad", "fmingrad_Rprop", "logit", "pop_codes",
           "unique_stable", "fp64_peak"]
