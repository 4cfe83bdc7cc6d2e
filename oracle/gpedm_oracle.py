"""CPU restatement (numpy / scipy-LAPACK, FP64) of GPEDM's GP fit + predict hot path.

*** TEST INFRASTRUCTURE - NOT PRODUCT CODE. ***
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import this module, and only as the checker / the timed
CPU baseline.  Nothing under `gpedm_b200/` imports it; the product path has no CPU
fallback and fails loudly when the CUDA library is missing.

Why a restatement: the reference is an R package and R is not installed in the build
image (no `R`, `Rscript`, `R.h`, `Rcpp.h`), so the reference
cannot be executed here.  Third-party arithmetic on the path that is NOT under
`/root/reference` (all unpinned in the reference's DESCRIPTION:26-29, no lockfile):
  * `laGP::distance`  (CRAN laGP) - pairwise squared Euclidean distance, restated in
    `distance()` below as difference-then-square accumulated over coordinates in order;
  * `Matrix::chol` / `Matrix::chol2inv` (CRAN Matrix) -> LAPACK `dpotrf` / `dpotri`,
    called here through `scipy.linalg.lapack` (the same routines);
  * base-R `%*%` -> BLAS `dgemm`/`dgemv`, here numpy `@`.
Parity pinning: the reference has no tests; the restatement is pinned against the
known answers printed in the reference README (K1-K5, see tests/test_oracle_readme.py),
which it reproduces to every printed digit.

Every function cites the reference lines it follows (paths relative to /root/reference).
"""
import math

import numpy as np
from scipy.linalg import lapack

VEMIN = SIGMA2MIN = 0.0001  # R/GPfunctions.R:591
VEMAX = SIGMA2MAX = 4.9999  # R/GPfunctions.R:592
RHOMIN, RHOMAX = 0.0001, 1.0  # R/GPfunctions.R:593


# --------------------------------------------------------------------------- helpers
def logit(x, lo=0.0, hi=1.0):
    """R/GPfunctions.R:1343-1345."""
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.log((np.asarray(x, dtype=np.float64) - lo) / (hi - np.asarray(x, dtype=np.float64)))


def unique_stable(v):
    """R's unique(): first-occurrence order."""
    seen, out = set(), []
    for e in v:
        if e not in seen:
            seen.add(e)
            out.append(e)
    return out


def _sd(v):
    v = np.asarray(v, dtype=np.float64)
    v = v[~np.isnan(v)]
    return float(np.std(v, ddof=1)) if v.size > 1 else float("nan")


def _mean(v):
    v = np.asarray(v, dtype=np.float64)
    v = v[~np.isnan(v)]
    return float(np.mean(v)) if v.size else float("nan")


def getR2(obs, pred):
    """R/GPfunctions.R:1366-1370."""
    obs = np.asarray(obs, dtype=np.float64)
    pred = np.asarray(pred, dtype=np.float64)
    ok = ~(np.isnan(obs) | np.isnan(pred))
    o, p = obs[ok], pred[ok]
    return 1.0 - np.sum((o - p) ** 2) / np.sum((o - np.mean(o)) ** 2)


def getR2pop(obs, pred, pop, type="within"):
    """R/GPfunctions.R:1394-1437."""
    obs = np.asarray(obs, dtype=np.float64)
    pred = np.asarray(pred, dtype=np.float64)
    pop = np.asarray(pop, dtype=object)
    up = unique_stable(pop)
    if type == "within":
        R2pop, rmsepop = {}, {}
        for u in up:
            ind = pop == u
            R2pop[u] = getR2(obs[ind], pred[ind])
            rmsepop[u] = math.sqrt(_mean((obs[ind] - pred[ind]) ** 2))
        return {"R2pop": R2pop, "rmsepop": rmsepop}
    obs_s, pred_s = obs.copy(), pred.copy()
    for u in up:
        ind = pop == u
        m = _mean(obs[ind])
        s = _sd(obs[ind]) if type == "scaled" else 1.0
        obs_s[ind] = (obs_s[ind] - m) / s
        pred_s[ind] = (pred_s[ind] - m) / s
    return getR2(obs_s, pred_s)


# ------------------------------------------------------------------ lag construction
def makelags_subrt(y, pop, E, tau, forecast=False):
    """R/GPfunctions.R:1824-1857.  y: (n, num_vars) matrix.  Returns lag matrix
    (columns var-major, then lag 1..E) or the forecast matrix."""
    y = np.asarray(y, dtype=np.float64)
    if y.ndim == 1:
        y = y[:, None]
    pop = np.asarray(pop, dtype=object)
    up = unique_stable(pop)
    n, nv = y.shape
    out = np.full((n, nv * E), np.nan)
    outfore = np.full((len(up) * tau, nv * E), np.nan)
    popfore = np.repeat(np.array(up, dtype=object), tau)
    for u in up:
        col = 0
        ind = np.where(pop == u)[0]
        indfore = np.where(popfore == u)[0]
        for j in range(nv):
            ts = y[ind, j]
            if len(ts) <= E * tau:
                raise ValueError("time series length <= E*tau for pop %s" % u)
            for i in range(1, E + 1):
                out[ind, col] = np.concatenate([np.full(tau * i, np.nan), ts[: len(ts) - tau * i]])
                outfore[indfore, col] = ts[len(ts) - tau * i: len(ts) - tau * (i - 1)]
                col += 1
    return outfore if forecast else out


# ------------------------------------------------------------------ numerical core
def distance(X1, X2=None):
    """laGP::distance (external C; call sites R/GPfunctions.R:541,787): squared Euclidean
    distance, D[i,j] = sum_k (X1[i,k]-X2[j,k])^2 accumulated over k in order."""
    X1 = np.asarray(X1, dtype=np.float64)
    if X1.ndim == 1:
        X1 = X1[:, None]
    X2 = X1 if X2 is None else np.asarray(X2, dtype=np.float64)
    if X2.ndim == 1:
        X2 = X2[:, None]
    D = np.zeros((X1.shape[0], X2.shape[0]))
    for k in range(X1.shape[1]):
        diff = X1[:, k][:, None] - X2[:, k][None, :]
        D += diff * diff
    return D


def getpriors(phi, ve, sigma2, rho, modeprior, sigma2max, vemax):
    """R/GPfunctions.R:752-777."""
    lam_phi = (2.0 ** (modeprior - 1)) ** 2 * math.pi / 2
    lp_phi = -0.5 * np.sum(phi ** 2) / lam_phi
    dlp_phi = -(phi ** 1) / lam_phi
    lp_sigma2 = math.log(sigma2 / sigma2max) + math.log(1 - sigma2 / sigma2max)
    dlp_sigma2 = 1.0 / sigma2 - 1.0 / (sigma2max - sigma2)
    lp_ve = math.log(ve / vemax) + math.log(1 - ve / vemax)
    dlp_ve = 1.0 / ve - 1.0 / (vemax - ve)
    lp_rho = math.log(rho) + math.log(1 - rho)
    dlp_rho = 1.0 / rho - 1.0 / (1 - rho)
    lp = lp_phi + lp_ve + lp_sigma2 + lp_rho
    dlp = np.concatenate([dlp_phi, [dlp_ve, dlp_sigma2, dlp_rho]])
    return lp, dlp


def getcov(phi, sigma2, rho, X, X2, pop, pop2, rhomat=None, rhomat_names=None):
    """R/GPfunctions.R:779-816.  Returns (popsame, Cd) with Cd of shape (nrow(X2), nrow(X)).
    `rhomat` is addressed by name via `rhomat_names` (R dimnames, :806)."""
    sq = np.sqrt(phi)
    lC0 = -distance(X2 * sq[None, :], X * sq[None, :])
    pop = np.asarray(pop, dtype=object)
    pop2 = np.asarray(pop2, dtype=object)
    popsame = (pop2[:, None] == pop[None, :]) * 1.0
    if rhomat is not None:
        up = unique_stable(pop)
        names = [str(n) for n in rhomat_names]
        Rmat = np.ones((X2.shape[0], X.shape[0]))
        for ui in up:
            for uj in up:
                indi = np.where(pop2 == ui)[0]
                indj = np.where(pop == uj)[0]
                Rmat[np.ix_(indi, indj)] = rhomat[names.index(str(ui)), names.index(str(uj))]
        Cd = sigma2 * np.exp(lC0) * Rmat
    else:
        Cd = sigma2 * np.exp(lC0) * (popsame + rho * (1 - popsame))
    return popsame, Cd


def getcovinv(Sigma):
    """R/GPfunctions.R:818-830: chol (LAPACK dpotrf, upper) + chol2inv (dpotri).
    Returns (L upper factor dense, iKVs dense symmetric)."""
    c, info = lapack.dpotrf(Sigma, lower=0, clean=1)
    if info != 0:
        raise np.linalg.LinAlgError("dpotrf info=%d (matrix not positive definite)" % info)
    inv, info = lapack.dpotri(c, lower=0)
    if info != 0:
        raise np.linalg.LinAlgError("dpotri info=%d" % info)
    iu = np.triu_indices_from(inv, 1)
    inv[(iu[1], iu[0])] = inv[iu]
    return c, inv


def getlikegrad(parst, Y, X, pop, modeprior, fixedpars, rhofixed, rhomatrix, D, returngradonly,
                rhomat_names=None):
    """R/GPfunctions.R:586-750.  `D` = list of per-coordinate squared-distance matrices."""
    parst = np.array(parst, dtype=np.float64)
    d = X.shape[1]
    with np.errstate(over="ignore"):
        phi = np.exp(parst[:d])
        ve = (VEMAX - VEMIN) / (1 + np.exp(-parst[d])) + VEMIN
        sigma2 = (SIGMA2MAX - SIGMA2MIN) / (1 + np.exp(-parst[d + 1])) + SIGMA2MIN
        rho = (RHOMAX - RHOMIN) / (1 + np.exp(-parst[d + 2])) + RHOMIN
    if len(unique_stable(pop)) == 1 or rhomatrix is not None:  # :600-602
        rhofixed = 0.5
    if rhofixed is not None:  # :605-608
        rho = rhofixed
        parst[d + 2] = logit(rho, RHOMIN, RHOMAX)
    fixedpars = np.asarray(fixedpars, dtype=np.float64)
    fx = ~np.isnan(fixedpars[:d])
    if fx.any():  # :609-612
        phi[fx] = fixedpars[:d][fx]
        with np.errstate(divide="ignore"):
            parst[:d] = np.log(phi)
    if not np.isnan(fixedpars[d]):  # :613-616
        ve = fixedpars[d]
        parst[d] = logit(ve, VEMIN, VEMAX)
    if not np.isnan(fixedpars[d + 1]):  # :617-620
        sigma2 = fixedpars[d + 1]
        parst[d + 1] = logit(sigma2, SIGMA2MIN, SIGMA2MAX)
    dpars = np.concatenate([phi, [  # :623-625
        (ve - VEMIN) * (1 - (ve - VEMIN) / (VEMAX - VEMIN)),
        (sigma2 - SIGMA2MIN) * (1 - (sigma2 - SIGMA2MIN) / (SIGMA2MAX - SIGMA2MIN)),
        (rho - RHOMIN) * (1 - (rho - RHOMIN) / (RHOMAX - RHOMIN))]])
    lp, dlp = getpriors(phi, ve, sigma2, rho, modeprior, SIGMA2MAX, VEMAX)  # :628
    popsame, Cd = getcov(phi, sigma2, rho, X, X, pop, pop, rhomatrix, rhomat_names)  # :633
    N = Cd.shape[0]
    Id = np.eye(N)
    Sigma = Cd + ve * Id  # :637
    Sigma_sym = (Sigma + Sigma.T) / 2  # :641
    L, iKVs = getcovinv(Sigma_sym)  # :642
    a = iKVs @ Y  # :647
    like = -0.5 * (Y @ a) - np.sum(np.log(np.diag(L)))  # :648
    dl = 0 * dlp
    vQ = 0.5 * (np.outer(a, a) - iKVs).ravel(order="F")  # :652
    for i in range(d):  # :659-665
        if np.isnan(fixedpars[i]):
            dC = -(D[i] * Cd)  # multMat, src/rcpp_la_utils.cpp:6-10
            dl[i] = 0.5 * np.dot(vQ, dC.ravel(order="F"))  # innerProd, :17-19
    if np.isnan(fixedpars[d]):  # :688-692
        dl[d] = np.dot(vQ, Id.ravel(order="F"))
    if np.isnan(fixedpars[d + 1]):  # :694-698
        dl[d + 1] = np.dot(vQ, (Cd / sigma2).ravel(order="F"))
    if rhofixed is not None:  # :701-709
        dl[d + 2] = 0
    else:
        dC = Cd * (1 - popsame) / rho
        dl[d + 2] = np.dot(vQ, dC.ravel(order="F"))
    J = dl + dlp  # :719
    neglgrad = -J * dpars
    neglpost = -(like + lp)
    if returngradonly:
        return {"nllpost": float(neglpost), "grad": neglgrad}
    SS = float(Y @ a)  # :728
    logdet = float(np.sum(np.log(np.diag(L))))
    pars = np.concatenate([phi, [ve, sigma2, rho]])
    mpt = Cd @ a  # :734
    Cdt = Cd - Cd @ iKVs @ Cd  # :735
    lnL_LOO = 0.5 * np.sum(np.log(np.diag(iKVs))) - 0.5 * np.sum(a ** 2 / np.diag(iKVs))  # :742
    df = float(np.sum(np.diag(Cd @ iKVs)))  # :743
    return {"pars": pars, "parst": parst, "nllpost": float(neglpost), "grad": neglgrad,
            "lnL_LOO": float(lnL_LOO), "df": df, "mean": mpt, "cov": Cdt,
            "covm": {"popsame": popsame, "Cd": Cd, "Sigma": Sigma, "iKVs": iKVs},
            "SS": SS, "logdet": logdet, "lp": float(lp), "a": a}


def fmingrad_Rprop(initparst, Y, X, pop, modeprior, fixedpars, rhofixed, rhomatrix,
                   rhomat_names=None, trace=None):
    """R/GPfunctions.R:522-584.  `trace` (optional list) receives (parst, nll, grad) per evaluation."""
    initparst = np.array(initparst, dtype=np.float64)
    p = len(initparst)
    Delta0 = 0.1 * np.ones(p)
    Deltamin, Deltamax = 1e-6, 50.0
    eta_minus = 0.5 - 1
    eta_plus = 1.2 - 1
    maxiter = 200
    D = [distance(X[:, i]) for i in range(X.shape[1])]  # :539-542
    out = getlikegrad(initparst, Y, X, pop, modeprior, fixedpars, rhofixed, rhomatrix, D, True, rhomat_names)
    nllpost, grad = out["nllpost"], out["grad"]
    if trace is not None:
        trace.append((initparst.copy(), nllpost, grad.copy()))
    s = math.sqrt(float(grad @ grad))
    pa = initparst
    it = 0
    dele = Delta0
    df = 10.0
    while s > 1e-4 and it < maxiter and df > 1e-7:  # :556
        panew = pa - np.sign(grad) * dele  # :559
        new = getlikegrad(panew, Y, X, pop, modeprior, fixedpars, rhofixed, rhomatrix, D, True, rhomat_names)
        nll_new, grad_new = new["nllpost"], new["grad"]
        if trace is not None:
            trace.append((panew.copy(), nll_new, grad_new.copy()))
        s = math.sqrt(float(grad_new @ grad_new))
        df = abs(nll_new / nllpost - 1)  # :565
        gc = grad * grad_new
        tdel = dele * (1 + eta_plus * (gc > 0) + eta_minus * (gc < 0))  # :569
        dele = np.where(tdel < Deltamin, Deltamin, np.where(tdel > Deltamax, Deltamax, tdel))
        pa, grad, nllpost = panew, grad_new, nll_new
        it += 1
    out = getlikegrad(pa, Y, X, pop, modeprior, fixedpars, rhofixed, rhomatrix, D, False, rhomat_names)  # :580
    out["iter"] = it
    return out


def getGPgrad(phi, Xnew, X, Cdi, iKVs, Y):
    """R/GPfunctions.R:1863-1865 (d >= 2; for d == 1 the reference's diag(phi) misbehaves,
    SURVEY appendix A17 - the analytic formula is used there)."""
    diff = Xnew[None, :] - X  # (N, d)
    w = np.ravel(Cdi) * (iKVs @ Y)
    return -2.0 * (diff * phi[None, :]).T @ w


# ------------------------------------------------------------------ prediction cores
def predict_newdata_core(phi, sigma2, rho, ve, X, Y, Pop, iKVs, Xnew, Popnew, rhomatrix=None,
                         rhomat_names=None, returnGPgrad=False):
    """R/GPfunctions.R:1012-1036."""
    _, Cs = getcov(phi, sigma2, rho, X, Xnew, Pop, Popnew, rhomatrix, rhomat_names)
    ymean = Cs @ (iKVs @ Y)
    yvar = np.array([sigma2 - Cs[i] @ iKVs @ Cs[i] for i in range(Xnew.shape[0])])
    grad = None
    if returnGPgrad:
        grad = np.vstack([getGPgrad(phi, Xnew[i], X, Cs[i], iKVs, Y) for i in range(Xnew.shape[0])])
    return ymean, yvar, grad


def predict_loo_core(sigma2, Cd, Sigma, Y, Pop, Time, Primary, exclradius=0, phi=None, X=None,
                     returnGPgrad=False):
    """R/GPfunctions.R:1056-1086 (brute force: one Cholesky + inverse per point)."""
    Pop = np.asarray(Pop, dtype=object)
    nd = int(np.sum(Primary))
    N = len(Y)
    ymean, yvar = np.zeros(nd), np.zeros(nd)
    grad = np.full((nd, X.shape[1]), np.nan) if returnGPgrad else None
    key = [str(p) + str(t) for p, t in zip(Pop, Time)]
    aug = np.where(~np.asarray(Primary))[0]
    for i in range(nd):
        excl = [j for j in range(i - exclradius, i + exclradius + 1) if 0 <= j < N and Pop[j] == Pop[i]]
        ek = {key[j] for j in excl}
        dups = [j for j in aug if key[j] in ek]
        keep = np.setdiff1d(np.arange(N), np.array(excl + dups, dtype=int))
        Cdi = Cd[i, keep]
        S = Sigma[np.ix_(keep, keep)]
        S = (S + S.T) / 2
        _, iK = getcovinv(S)
        ymean[i] = Cdi @ iK @ Y[keep]
        yvar[i] = sigma2 - Cdi @ iK @ Cdi
        if returnGPgrad:
            grad[i] = getGPgrad(phi, X[i], X[keep], Cdi, iK, Y[keep])
    return ymean, yvar, grad


def predict_lto_core(sigma2, Cd, Sigma, Y, Time, Primary, exclradius=0, phi=None, X=None,
                     returnGPgrad=False):
    """R/GPfunctions.R:1140-1176."""
    Time = np.asarray(Time)
    timevals = unique_stable(Time.tolist())
    nt = len(timevals)
    nd = int(np.sum(Primary))
    ymean, yvar = np.full(nd, np.nan), np.full(nd, np.nan)
    grad = np.full((nd, X.shape[1]), np.nan) if returnGPgrad else None
    Tp = Time[np.asarray(Primary)]
    for i in range(nt):
        excl = [timevals[j] for j in range(i - exclradius, i + exclradius + 1) if 0 <= j < nt]
        ind = np.where(Tp == timevals[i])[0]
        train = np.where(~np.isin(Time, excl))[0]
        Cdi = Cd[np.ix_(ind, train)]
        S = Sigma[np.ix_(train, train)]
        S = (S + S.T) / 2
        _, iK = getcovinv(S)
        ymean[ind] = Cdi @ iK @ Y[train]
        yvar[ind] = np.diag(sigma2 - Cdi @ iK @ Cdi.T)
        if returnGPgrad:
            for j in ind:
                grad[j] = getGPgrad(phi, X[j], X[train], Cd[j, train], iK, Y[train])
    return ymean, yvar, grad


def predict_sequential_core(sigma2, ve, Cd, Y, Pop, Time):
    """R/GPfunctions.R:1196-1245 (returns ymean and ysd2 over all N rows; note the quirk that
    never-updated entries carry (sigma2+ve), a variance, in the sd slot, :1216,1236)."""
    Pop = np.asarray(Pop, dtype=object)
    Time = np.asarray(Time)
    up = unique_stable(Pop)
    timevals = unique_stable(Time.tolist())
    nt = len(timevals)
    N = len(Y)
    S = Cd.copy()
    ymean1 = np.zeros((N, nt))
    ysd1 = (sigma2 + ve) * np.ones((N, nt))
    for i in range(nt - 1):
        ind = np.where(Time == timevals[i])[0]
        SS = S[np.ix_(ind, ind)] + ve * np.eye(len(ind))
        SS = (SS + SS.T) / 2
        _, iK = getcovinv(SS)
        k = iK @ (Y[ind] - ymean1[ind, i])
        ymean1[:, i + 1] = ymean1[:, i] + S[:, ind] @ k
        S = S - S[:, ind] @ iK @ S[ind, :]
        ysd1[:, i + 1] = np.sqrt(np.abs(np.diag(S)) + ve)
    ymean = np.full(N, np.nan)
    ysd2 = np.full(N, sigma2 + ve)
    for i in range(nt):
        for u in up:
            ind = np.where((Time == timevals[i]) & (Pop == u))[0]
            if len(ind):
                ymean[ind] = ymean1[ind, i]
                ysd2[ind] = ysd1[ind, i]
    return ymean, ysd2


# ------------------------------------------------------------------ fitGP / predict front ends
def _col(data, key):
    return np.asarray(data[key])


def _as_matrix(data, keys):
    if isinstance(keys, str):
        keys = [keys]
    return np.column_stack([np.asarray(data[k], dtype=np.float64) for k in keys])


def fitGP(data=None, y=None, x=None, pop=None, time=None, E=None, tau=None, scaling="global",
          initpars=None, modeprior=1, fixedpars=None, rhofixed=None, rhomatrix=None,
          rhomatrix_names=None, augdata=None, predictmethod=None, newdata=None, xnew=None,
          popnew=None, timenew=None, ynew=None, returnGPgrad=False, exclradius=0, _prep_only=False):
    """R/GPfunctions.R:221-520 (linprior not restated: it only pre-processes y with lm())."""
    if (E is None) != (tau is None):
        raise ValueError("Both E and tau must be supplied if generating lags internally.")
    if x is None and (E is None or tau is None):
        raise ValueError("x and/or E and tau must be supplied")
    if rhofixed is not None:
        if rhomatrix is not None:
            raise ValueError("Supply either rhofixed or rhomatrix, not both")
        rhofixed = min(max(rhofixed, 0.0001), 0.9999)  # :241-248
    inputs = {}
    if data is not None:  # :261-276
        inputs["y_names"] = y
        yv = np.asarray(data[y], dtype=np.float64)
        if x is not None:
            inputs["x_names"] = [x] if isinstance(x, str) else list(x)
            xv = _as_matrix(data, x)
        else:
            xv = None
        if pop is not None:
            inputs["pop_names"] = pop
            popv = np.asarray(data[pop], dtype=object)
        else:
            popv = None
        if time is not None:
            inputs["time_names"] = time
            timev = np.asarray(data[time], dtype=np.float64)
        else:
            timev = None
    else:
        yv = np.asarray(y, dtype=np.float64)
        xv = None if x is None else np.asarray(x, dtype=np.float64)
        popv = None if pop is None else np.asarray(pop, dtype=object)
        timev = None if time is None else np.asarray(time, dtype=np.float64)
    n = len(yv)
    if popv is None:
        popv = np.array([1] * n, dtype=object)
    if timev is None:  # :283-289
        timev = np.zeros(n)
        for u in unique_stable(popv):
            m = popv == u
            timev[m] = np.arange(1, m.sum() + 1)
    if E is not None:  # :292-306
        if xv is None:
            xv = yv
        xv = makelags_subrt(xv, popv, E, tau)
        inputs["E"], inputs["tau"] = E, tau
    if xv.ndim == 1:
        xv = xv[:, None]
    d = xv.shape[1]
    if augdata is not None:  # :313-330
        xaug = _as_matrix(augdata, inputs["x_names"])
        yaug = np.asarray(augdata[inputs["y_names"]], dtype=np.float64)
        timeaug = np.asarray(augdata[inputs["time_names"]], dtype=np.float64)
        popaug = (np.array([1] * len(yaug), dtype=object) if "pop_names" not in inputs
                  else np.asarray(augdata[inputs["pop_names"]], dtype=object))
        yd = np.concatenate([yv, yaug])
        xd = np.vstack([xv, xaug])
        timed = np.concatenate([timev, timeaug])
        popd = np.concatenate([popv, popaug])
        primary = np.concatenate([np.ones(n, bool), np.zeros(len(yaug), bool)])
    else:
        yd, xd, timed, popd, primary = yv, xv, timev, popv, np.ones(n, bool)
    up = unique_stable(popd)
    # scaling :359-405
    if scaling == "none":
        ymeans, ysds = _mean(yd), _sd(yd)
        xmeans = np.array([_mean(xd[:, j]) for j in range(d)])
        xsds = np.array([_sd(xd[:, j]) for j in range(d)])
        yds, xds = yd.copy(), xd.copy()
    elif scaling == "global":
        ymeans, ysds = _mean(yd), _sd(yd)
        xmeans = np.array([_mean(xd[:, j]) for j in range(d)])
        xsds = np.array([_sd(xd[:, j]) for j in range(d)])
        yds = (yd - ymeans) / ysds
        xds = (xd - xmeans[None, :]) / xsds[None, :]
    elif scaling == "local":
        ymeans = {u: _mean(yd[popd == u]) for u in up}
        ysds = {u: _sd(yd[popd == u]) for u in up}
        xmeans = {u: np.array([_mean(xd[popd == u, j]) for j in range(d)]) for u in up}
        xsds = {u: np.array([_sd(xd[popd == u, j]) for j in range(d)]) for u in up}
        yds, xds = yd.copy(), xd.copy()
        for u in up:
            m = popd == u
            yds[m] = (yd[m] - ymeans[u]) / ysds[u]
            xds[m] = (xd[m] - xmeans[u][None, :]) / xsds[u][None, :]
    else:
        raise ValueError("scaling must be one of global, local, none")
    if _prep_only:  # the section getrhomatrix duplicates from fitGP (R/rhomatrix.R:26-160)
        return dict(yds=yds, xds=xds, popd=popd, timed=timed, primary=primary, up=up, d=d)
    if initpars is None:  # :408-410
        initpars = np.concatenate([np.full(d, 0.1), [0.001, 1 - 0.001, 0.5]])
    initpars = np.array(initpars, dtype=np.float64)
    if fixedpars is None:
        fixedpars = np.full(d + 2, np.nan)
    fixedpars = np.asarray(fixedpars, dtype=np.float64)
    if len(fixedpars) != d + 2:
        raise ValueError("Length of fixedpars must be number of predictors + 2 (phi, ve, sigma2)")
    fi = ~np.isnan(fixedpars)
    initpars[:d + 2][fi] = fixedpars[fi]
    with np.errstate(divide="ignore"):
        initparst = np.concatenate([np.log(initpars[:d]), [logit(initpars[d], VEMIN, VEMAX),  # :425-426
                                                           logit(initpars[d + 1], SIGMA2MIN, SIGMA2MAX),
                                                           logit(initpars[d + 2], 0.0, 1.0)]])
    if len(up) == 1:
        initparst[d + 2] = 0  # :429
    if rhomatrix is not None:  # :432-443
        rhomatrix = np.asarray(rhomatrix, dtype=np.float64)
        if rhomatrix_names is None:
            rhomatrix_names = up
        if not all(str(u) in [str(nm) for nm in rhomatrix_names] for u in up):
            raise ValueError("Population names are not all present in rhomatrix.")
        inputs["rhomatrix"] = rhomatrix
        inputs["rhomatrix_names"] = list(rhomatrix_names)
    completerows = ~(np.isnan(yds) | np.isnan(xds).any(axis=1))  # :446
    Y, X = yds[completerows], xds[completerows]
    Pop, Time, Primary = popd[completerows], timed[completerows], primary[completerows]
    output = fmingrad_Rprop(initparst, Y, X, Pop, modeprior, fixedpars, rhofixed, rhomatrix, rhomatrix_names)
    ve = output["pars"][d]
    ypred = np.full(len(yd), np.nan)
    yfsd = np.full(len(yd), np.nan)
    ysd = np.full(len(yd), np.nan)
    ypred[completerows] = output["mean"]  # :458-461
    with np.errstate(invalid="ignore"):
        yfsd[completerows] = np.sqrt(np.diag(output["cov"]))
        ysd[completerows] = np.sqrt(np.diag(output["cov"]) + ve)
    ypred, yfsd, ysd = ypred[primary], yfsd[primary], ysd[primary]
    if scaling == "global":  # :469-482
        ypred = ypred * ysds + ymeans
        yfsd = np.sqrt(yfsd ** 2 * ysds ** 2)
        ysd = np.sqrt(ysd ** 2 * ysds ** 2)
    if scaling == "local":
        for u in up:
            m = popv == u
            ypred[m] = ypred[m] * ysds[u] + ymeans[u]
            yfsd[m] = np.sqrt(yfsd[m] ** 2 * ysds[u] ** 2)
            ysd[m] = np.sqrt(ysd[m] ** 2 * ysds[u] ** 2)
    inputs.update(dict(y=yv, x=xv, yd=yd, xd=xd, pop=popv, time=timev, completerows=completerows,
                       Y=Y, X=X, Pop=Pop, Time=Time, Primary=Primary))
    output["inputs"] = inputs
    output["scaling"] = dict(scaling=scaling, ymeans=ymeans, ysds=ysds, xmeans=xmeans, xsds=xsds)
    output["rhofixed"] = rhofixed
    output["insampresults"] = dict(timestep=timev, pop=popv, predmean=ypred, predfsd=yfsd, predsd=ysd, obs=yv)
    stats = dict(R2=getR2(yv, ypred), rmse=math.sqrt(_mean((yv - ypred) ** 2)), ln_post=-output["nllpost"],
                 ln_prior=output["lp"], lnL_LOO=output["lnL_LOO"], df=output["df"], SS=output["SS"],
                 logdet=output["logdet"])
    if len(unique_stable(popv)) > 1:  # :502-507
        output["insampfitstatspop"] = getR2pop(yv, ypred, popv)
        stats["R2centered"] = getR2pop(yv, ypred, popv, "centered")
        stats["R2scaled"] = getR2pop(yv, ypred, popv, "scaled")
    output["insampfitstats"] = stats
    output["fit_args"] = dict(y=y, x=x, pop=pop, time=time, E=E, tau=tau, scaling=scaling, modeprior=modeprior,
                              fixedpars=None if not fi.any() else fixedpars, rhofixed=rhofixed,
                              rhomatrix=rhomatrix, rhomatrix_names=rhomatrix_names)
    if predictmethod is not None or xnew is not None or newdata is not None:  # :512-515
        output.update(predict(output, predictmethod, newdata, xnew, popnew, timenew, ynew, returnGPgrad, exclradius))
    return output


def predict(obj, predictmethod="loo", newdata=None, xnew=None, popnew=None, timenew=None, ynew=None,
            returnGPgrad=False, exclradius=0):
    """R/GPfunctions.R:880-1341 (fisheries and linprior branches not restated)."""
    pars = obj["pars"]
    inp, sc = obj["inputs"], obj["scaling"]
    X, Y, Pop, yv = inp["X"], inp["Y"], inp["Pop"], inp["y"]
    d = X.shape[1]
    phi, ve, sigma2, rho = pars[:d], pars[d], pars[d + 1], pars[d + 2]
    iKVs = obj["covm"]["iKVs"]
    rhomatrix, rnames = inp.get("rhomatrix"), inp.get("rhomatrix_names")
    scaling, ymeans, ysds, xmeans, xsds = sc["scaling"], sc["ymeans"], sc["ysds"], sc["xmeans"], sc["xsds"]
    gpgrad = None
    if newdata is not None or xnew is not None:
        if newdata is not None:  # :933-943
            yn = inp.get("y_names")
            if isinstance(yn, str) and yn in newdata:
                ynew = np.asarray(newdata[yn], dtype=np.float64)
            if inp.get("x_names") is not None:
                xnew = _as_matrix(newdata, inp["x_names"])
            if inp.get("pop_names") is not None:
                popnew = np.asarray(newdata[inp["pop_names"]], dtype=object)
            if inp.get("time_names") is not None:
                timenew = np.asarray(newdata[inp["time_names"]], dtype=np.float64)
        Tslp = len(ynew) if ynew is not None else np.atleast_2d(np.asarray(xnew).T).T.shape[0]
        if popnew is None:
            popnew = np.array([1] * Tslp, dtype=object)
        popnew = np.asarray(popnew, dtype=object)
        if timenew is None:
            timenew = np.zeros(Tslp)
            for u in unique_stable(popnew):
                m = popnew == u
                timenew[m] = np.arange(1, m.sum() + 1)
        if inp.get("E") is not None:  # :966-971
            if xnew is None:
                xnew = ynew
            xnew = makelags_subrt(xnew, popnew, inp["E"], inp["tau"])
        xnew = np.asarray(xnew, dtype=np.float64)
        if xnew.ndim == 1:
            xnew = xnew[:, None]
        xnews = xnew.copy()
        if scaling == "global":
            xnews = (xnew - xmeans[None, :]) / xsds[None, :]
        if scaling == "local":
            for u in unique_stable(popnew):
                m = popnew == u
                xnews[m] = (xnew[m] - xmeans[u][None, :]) / xsds[u][None, :]
        comp = ~np.isnan(xnews).any(axis=1)
        ypred = np.full(Tslp, np.nan)
        yfsd = np.full(Tslp, np.nan)
        ysd = np.full(Tslp, np.nan)
        if comp.any():
            Xn, Popn = xnews[comp], popnew[comp]
            ymean, yvar, G = predict_newdata_core(phi, sigma2, rho, ve, X, Y, Pop, iKVs, Xn, Popn, rhomatrix,
                                                  rnames, returnGPgrad)
            ypred[comp] = ymean
            with np.errstate(invalid="ignore"):
                yfsd[comp] = np.sqrt(yvar)
                ysd[comp] = np.sqrt(yvar + ve)
            if returnGPgrad:
                gpgrad = np.full(xnew.shape, np.nan)
                gpgrad[comp] = G
    else:
        if predictmethod is None:
            predictmethod = "loo"
        Cd, Sigma = obj["covm"]["Cd"], obj["covm"]["Sigma"]
        Time, Primary = inp["Time"], inp["Primary"]
        cr = inp["completerows"][:len(yv)]
        ypred = np.full(len(yv), np.nan)
        yfsd = np.full(len(yv), np.nan)
        ysd = np.full(len(yv), np.nan)
        if predictmethod in ("loo", "lto"):
            if predictmethod == "loo":
                ymean, yvar, G = predict_loo_core(sigma2, Cd, Sigma, Y, Pop, Time, Primary, exclradius, phi, X,
                                                  returnGPgrad)
            else:
                ymean, yvar, G = predict_lto_core(sigma2, Cd, Sigma, Y, Time, Primary, exclradius, phi, X,
                                                  returnGPgrad)
            ypred[cr] = ymean
            with np.errstate(invalid="ignore"):
                yfsd[cr] = np.sqrt(yvar)
                ysd[cr] = np.sqrt(yvar + ve)
            if returnGPgrad:
                gpgrad = np.full((len(yv), d), np.nan)
                gpgrad[cr] = G
        elif predictmethod == "sequential":
            ymean, ysd2 = predict_sequential_core(sigma2, ve, Cd, Y, Pop, Time)
            ypred[cr] = ymean[Primary]
            with np.errstate(invalid="ignore"):
                yfsd[cr] = np.sqrt(ysd2 ** 2 - ve)[Primary]
            ysd[cr] = ysd2[Primary]
        else:
            raise ValueError("predictmethod must be loo, lto or sequential")
        popnew, timenew, ynew = inp["pop"], inp["time"], yv
    # unscale :1260-1287
    if scaling == "global":
        ypred = ypred * ysds + ymeans
        yfsd = np.sqrt(yfsd ** 2 * ysds ** 2)
        ysd = np.sqrt(ysd ** 2 * ysds ** 2)
        if gpgrad is not None:
            gpgrad = gpgrad * ysds / xsds[None, :]
    if scaling == "local":
        for u in unique_stable(popnew):
            m = popnew == u
            ypred[m] = ypred[m] * ysds[u] + ymeans[u]
            yfsd[m] = np.sqrt(yfsd[m] ** 2 * ysds[u] ** 2)
            ysd[m] = np.sqrt(ysd[m] ** 2 * ysds[u] ** 2)
            if gpgrad is not None:
                gpgrad[m] = gpgrad[m] * ysds[u] / xsds[u][None, :]
    out = {"outsampresults": dict(timestep=timenew, pop=popnew, predmean=ypred, predfsd=yfsd, predsd=ysd)}
    if ynew is not None:
        out["outsampresults"]["obs"] = ynew
        out["outsampfitstats"] = dict(R2=getR2(ynew, ypred), rmse=math.sqrt(_mean((ynew - ypred) ** 2)))
        if len(unique_stable(popnew)) > 1:
            out["outsampfitstatspop"] = getR2pop(ynew, ypred, popnew)
            out["outsampfitstats"]["R2centered"] = getR2pop(ynew, ypred, popnew, "centered")
            out["outsampfitstats"]["R2scaled"] = getR2pop(ynew, ypred, popnew, "scaled")
    if gpgrad is not None:
        out["GPgrad"] = gpgrad
    return out


def _rbind(a, b):
    return {k: np.concatenate([np.asarray(a[k]), np.asarray(b[k])]) for k in a}


def _rows(df, mask):
    return {k: np.asarray(v)[mask] for k, v in df.items()}


def predict_seq(obj, data, newdata, restart_initpars=False):
    """R/predict_seq.R:38-99: predict next time value, append it, refit warm-started via update()."""
    inp = obj["inputs"]
    if inp.get("time_names") is None:
        raise ValueError("Model must include `data` and `time` to use this function.")
    if inp.get("E") is not None:
        raise ValueError("Lags must be pre-generated to use this function. See option A1 in help(fitGP).")
    tn = inp["time_names"]
    newtimes = unique_stable(np.asarray(newdata[tn]).tolist())
    model, dataup, preds = obj, data, []
    fa = obj["fit_args"]
    for t in newtimes:
        nd = _rows(newdata, np.asarray(newdata[tn]) == t)
        preds.append(predict(model, newdata=nd)["outsampresults"])
        dataup = _rbind(dataup, nd)
        inits = None if restart_initpars else model["pars"]
        model = fitGP(data=dataup, initpars=inits, **fa)
    res = preds[0]
    for p in preds[1:]:
        res = _rbind(res, p)
    model["outsampresults"] = res
    model["outsampfitstats"] = dict(R2=getR2(res["obs"], res["predmean"]),
                                    rmse=math.sqrt(_mean((res["obs"] - res["predmean"]) ** 2)))
    return model


def predict_iter(obj, newdata, xlags=None):
    """R/predict_iter.R:57-163 (non-fisheries): iterated multi-step forecasts; the prediction at one
    time step becomes the first lag of the next, the other lags shift."""
    inp = obj["inputs"]
    if inp.get("time_names") is None:
        raise ValueError("Model must include `data` and `time` to use this function.")
    if inp.get("E") is not None:
        raise ValueError("Lags must be pre-generated to use this function. See option A1 in help(fitGP).")
    tn = inp["time_names"]
    if tn not in newdata:
        raise ValueError("Time column '%s' not found in newdata." % tn)
    if xlags is None:
        xlags = list(inp["x_names"])
    nd = {k: np.array(v, dtype=object if np.asarray(v).dtype == object else np.float64) for k, v in newdata.items()}
    pn = inp.get("pop_names")
    if pn is None:
        pn = "pop"
        nd[pn] = np.array([1] * len(nd[tn]), dtype=object)
    elif pn not in nd:
        raise ValueError("Pop column '%s' not found in newdata." % pn)
    newtimes = unique_stable(nd[tn].tolist())
    up = unique_stable(nd[pn])
    preds = []
    for i, t in enumerate(newtimes):
        mask = nd[tn] == t
        cur = _rows(nd, mask)
        if inp.get("pop_names") is None:
            cur = {k: v for k, v in cur.items() if k != pn}
        pr = predict(obj, newdata=cur)["outsampresults"]
        preds.append(pr)
        if np.any(np.isinf(pr["predmean"])):
            break
        if i + 1 < len(newtimes):
            nxt = nd[tn] == newtimes[i + 1]
            for u in up:
                pj = pr["predmean"][np.asarray(pr["pop"], dtype=object) == u] if inp.get("pop_names") is not None else pr["predmean"]
                sel_n = nxt & (nd[pn] == u)
                sel_c = mask & (nd[pn] == u)
                old = [nd[c][sel_c].copy() for c in xlags]
                nd[xlags[0]][sel_n] = pj
                for q in range(1, len(xlags)):
                    nd[xlags[q]][sel_n] = old[q - 1]
    res = preds[0]
    for p_ in preds[1:]:
        res = _rbind(res, p_)
    out = {"outsampresults": res}
    if "obs" in res and not np.all(np.isnan(np.asarray(res["obs"], dtype=float))):
        out["outsampfitstats"] = dict(R2=getR2(res["obs"], res["predmean"]),
                                      rmse=math.sqrt(_mean((np.asarray(res["obs"], float) - res["predmean"]) ** 2)))
    return out


def getconditionals(fit, xrange="default", extrap=0.01, nvals=25):
    """R/PlotsConditionals.R:166-266: conditional responses - for every population and predictor, nvals
    points along that predictor with all other (scaled) predictors held at 0.  Returns one dict per
    population: poppred, xval, predmean, predsd (each nvals x d), unscaled like the reference.
    Quirk kept: with xrange global the upper end uses max(X[]) over the whole matrix (:205)."""
    pars = fit["pars"]
    inp, sc = fit["inputs"], fit["scaling"]
    X, Y, Pop = inp["X"], inp["Y"], np.asarray(inp["Pop"], dtype=object)
    d = X.shape[1]
    phi, ve, sigma2, rho = pars[:d], pars[d], pars[d + 1], pars[d + 2]
    iKVs = fit["covm"]["iKVs"]
    scaling = sc["scaling"]
    xr = ("local" if scaling == "local" else "global") if xrange == "default" else xrange
    out = []
    for u in unique_stable(Pop):
        indi = Pop == u
        xval, pm, ps = np.zeros((nvals, d)), np.zeros((nvals, d)), np.zeros((nvals, d))
        poppred = np.array([u] * nvals, dtype=object)
        for i in range(d):
            if xr == "local":
                ext = (X[indi, i].max() - X[indi, i].min()) * extrap
                xgi = np.linspace(X[indi, i].min() - ext, X[indi, i].max() + ext, nvals)
            else:
                ext = (X[:, i].max() - X[:, i].min()) * extrap
                xgi = np.linspace(X[:, i].min() - ext, X.max() + ext, nvals)
            xp = np.zeros((nvals, d))
            xp[:, i] = xgi
            m, v, _ = predict_newdata_core(phi, sigma2, rho, ve, X, Y, Pop, iKVs, xp, poppred, inp.get("rhomatrix"),
                                           inp.get("rhomatrix_names"))
            pm[:, i], ps[:, i], xval[:, i] = m, np.sqrt(v), xgi
        if scaling == "global":
            pm = pm * sc["ysds"] + sc["ymeans"]
            ps = np.sqrt(ps ** 2 * sc["ysds"] ** 2)
            xval = xval * sc["xsds"][None, :] + sc["xmeans"][None, :]
        if scaling == "local":
            pm = pm * sc["ysds"][u] + sc["ymeans"][u]
            ps = np.sqrt(ps ** 2 * sc["ysds"][u] ** 2)
            xval = xval * sc["xsds"][u][None, :] + sc["xmeans"][u][None, :]
        out.append(dict(poppred=poppred, xval=xval, predmean=pm, predsd=ps))
    return out


# ------------------------------------------------------------------ extended precision
def chol_lower_ld(S):
    """Cholesky (lower) in x87 80-bit long double; O(N) python steps of vectorised numpy."""
    A = np.array(S, dtype=np.longdouble)
    n = A.shape[0]
    for j in range(n):
        A[j, j] = np.sqrt(A[j, j] - np.dot(A[j, :j], A[j, :j]))
        if j + 1 < n:
            A[j + 1:, j] = (A[j + 1:, j] - A[j + 1:, :j] @ A[j, :j]) / A[j, j]
    return np.tril(A)


def solve_lower_ld(L, B):
    """Forward substitution L Z = B in long double."""
    L = np.asarray(L, dtype=np.longdouble)
    Z = np.array(B, dtype=np.longdouble)
    n = L.shape[0]
    for i in range(n):
        Z[i] = (Z[i] - L[i, :i] @ Z[:i]) / L[i, i]
    return Z


def solve_upper_ld(U, B):
    U = np.asarray(U, dtype=np.longdouble)
    Z = np.array(B, dtype=np.longdouble)
    n = U.shape[0]
    for i in range(n - 1, -1, -1):
        Z[i] = (Z[i] - U[i, i + 1:] @ Z[i + 1:]) / U[i, i]
    return Z


def getcov_ld(phi, sigma2, rho, X, X2, pop, pop2, rhomat=None, rhomat_names=None):
    """getcov (R/GPfunctions.R:779-816) evaluated in long double."""
    ld = np.longdouble
    sq = np.sqrt(np.asarray(phi, dtype=ld))
    A = np.asarray(X2, dtype=ld) * sq[None, :]
    B = np.asarray(X, dtype=ld) * sq[None, :]
    D = np.zeros((A.shape[0], B.shape[0]), dtype=ld)
    for k in range(A.shape[1]):
        diff = A[:, k][:, None] - B[:, k][None, :]
        D += diff * diff
    pop = np.asarray(pop, dtype=object)
    pop2 = np.asarray(pop2, dtype=object)
    same = (pop2[:, None] == pop[None, :])
    if rhomat is not None:
        names = [str(n) for n in rhomat_names]
        i2 = np.array([names.index(str(p)) for p in pop2])
        i1 = np.array([names.index(str(p)) for p in pop])
        R = np.asarray(rhomat, dtype=ld)[np.ix_(i2, i1)]
        R = np.where(same, ld(1), R)
    else:
        R = np.where(same, ld(1), ld(rho))
    return ld(sigma2) * np.exp(-D) * R


def predict_newdata_ld(phi, sigma2, rho, ve, X, Y, Pop, Xnew, Popnew, rhomat=None, rhomat_names=None):
    """Extended-precision truth for R/GPfunctions.R:1012-1021 (mean and variance), via
    long-double Cholesky + triangular solves (no explicit inverse)."""
    ld = np.longdouble
    Cd = getcov_ld(phi, sigma2, rho, X, X, Pop, Pop, rhomat, rhomat_names)
    Sigma = Cd + ld(ve) * np.eye(len(Y), dtype=ld)
    L = chol_lower_ld(Sigma)
    Cs = getcov_ld(phi, sigma2, rho, X, Xnew, Pop, Popnew, rhomat, rhomat_names)
    z = solve_lower_ld(L, np.asarray(Y, dtype=ld))
    V = solve_lower_ld(L, Cs.T)  # N x M
    mean = V.T @ z
    var = ld(sigma2) - np.sum(V * V, axis=0)
    return mean, var


def loglik_ld(phi, sigma2, rho, ve, X, Y, Pop, rhomat=None, rhomat_names=None):
    """Extended-precision `like`, a = Sigma^-1 Y and diag(Sigma^-1) (R/GPfunctions.R:647-648, 742)."""
    ld = np.longdouble
    Cd = getcov_ld(phi, sigma2, rho, X, X, Pop, Pop, rhomat, rhomat_names)
    N = len(Y)
    Sigma = Cd + ld(ve) * np.eye(N, dtype=ld)
    L = chol_lower_ld(Sigma)
    z = solve_lower_ld(L, np.asarray(Y, dtype=ld))
    a = solve_upper_ld(L.T, z)
    like = -ld(0.5) * (z @ z) - np.sum(np.log(np.diag(L)))
    Linv = solve_lower_ld(L, np.eye(N, dtype=ld))
    dinv = np.sum(Linv * Linv, axis=0)
    return like, a, dinv


def _condition_ld(Sigma, Cd, Y, sigma2, focal, train):
    """Extended-precision conditional of rows `focal` on rows `train`:
    mean = Cd[f,t] Sigma[t,t]^-1 Y[t],  var = sigma2 - diag(Cd[f,t] Sigma[t,t]^-1 Cd[t,f])  via a
    long-double Cholesky and triangular solves (no explicit inverse, no cancellation-prone products)."""
    ld = np.longdouble
    if len(train) == 0:
        return np.zeros(len(focal), dtype=ld), np.full(len(focal), ld(sigma2))
    L = chol_lower_ld(Sigma[np.ix_(train, train)])
    z = solve_lower_ld(L, np.asarray(Y, dtype=ld)[train])
    V = solve_lower_ld(L, Cd[np.ix_(train, focal)])
    return V.T @ z, ld(sigma2) - np.sum(V * V, axis=0)


def predict_leaveout_ld(method, phi, sigma2, rho, ve, X, Y, Pop, Time, Primary, exclradius=0, rhomat=None,
                        rhomat_names=None):
    """Extended-precision truth for the loo (R/GPfunctions.R:1056-1086) and lto (:1140-1176) loops:
    the same excluded sets, every refactorisation done in 80-bit.  Returns (ymean, yvar) as long double."""
    ld = np.longdouble
    Pop = np.asarray(Pop, dtype=object)
    Time = np.asarray(Time)
    Primary = np.asarray(Primary, dtype=bool)
    N = len(Y)
    Cd = getcov_ld(phi, sigma2, rho, X, X, Pop, Pop, rhomat, rhomat_names)
    Sigma = Cd + ld(ve) * np.eye(N, dtype=ld)
    nd = int(np.sum(Primary))
    ymean, yvar = np.full(nd, np.nan, dtype=ld), np.full(nd, np.nan, dtype=ld)
    if method == "loo":
        key = [str(p) + str(t) for p, t in zip(Pop, Time)]
        aug = np.where(~Primary)[0]
        for i in range(nd):
            excl = [j for j in range(i - exclradius, i + exclradius + 1) if 0 <= j < N and Pop[j] == Pop[i]]
            ek = {key[j] for j in excl}
            dups = [j for j in aug if key[j] in ek]
            keep = np.setdiff1d(np.arange(N), np.array(excl + dups, dtype=int))
            m, v = _condition_ld(Sigma, Cd, Y, sigma2, np.array([i]), keep)
            ymean[i], yvar[i] = m[0], v[0]
    else:
        timevals = unique_stable(Time.tolist())
        nt = len(timevals)
        Tp = Time[Primary]
        for i in range(nt):
            excl = [timevals[j] for j in range(i - exclradius, i + exclradius + 1) if 0 <= j < nt]
            ind = np.where(Tp == timevals[i])[0]
            train = np.where(~np.isin(Time, excl))[0]
            ymean[ind], yvar[ind] = _condition_ld(Sigma, Cd, Y, sigma2, ind, train)
    return ymean, yvar


def predict_sequential_ld(phi, sigma2, rho, ve, X, Y, Pop, Time, rhomat=None, rhomat_names=None):
    """Extended-precision truth for the sequential loop (R/GPfunctions.R:1196-1245): the rows of time
    value i conditioned on every row of an EARLIER time value (what the successive Schur downdates of
    :1228-1233 amount to), each conditioning done from scratch in 80-bit.  Returns (ymean, ysd2) with the
    reference's first-time-step quirk (ymean 0, ysd = sigma2 + ve, :1216, :1236)."""
    ld = np.longdouble
    Pop = np.asarray(Pop, dtype=object)
    Time = np.asarray(Time)
    N = len(Y)
    Cd = getcov_ld(phi, sigma2, rho, X, X, Pop, Pop, rhomat, rhomat_names)
    Sigma = Cd + ld(ve) * np.eye(N, dtype=ld)
    timevals = unique_stable(Time.tolist())
    ymean = np.zeros(N, dtype=ld)
    ysd2 = np.full(N, ld(sigma2) + ld(ve))
    for i in range(1, len(timevals)):
        ind = np.where(Time == timevals[i])[0]
        train = np.where(np.isin(Time, timevals[:i]))[0]
        m, v = _condition_ld(Sigma, Cd, Y, sigma2, ind, train)
        ymean[ind] = m
        ysd2[ind] = np.sqrt(np.abs(v) + ld(ve))
    return ymean, ysd2


def getrhomatrix(data=None, y=None, x=None, pop=None, time=None, E=None, tau=None, scaling="global",
                 initpars=None, modeprior=1, fixedpars=None, augdata=None):
    """R/rhomatrix.R:22-186: the data preparation of fitGP, then one two-population fit per pair of
    populations (scaling "none" on the already scaled data, :169-182); rho of pair (i, j) fills a
    symmetric matrix with unit diagonal.  Returns (rhomatrix, names)."""
    pr = fitGP(data=data, y=y, x=x, pop=pop, time=time, E=E, tau=tau, scaling=scaling, augdata=augdata,
               _prep_only=True)
    up = pr["up"]
    n_pop = len(up)
    if n_pop == 1:
        raise ValueError("There must be more than 1 population to get a pairwise matrix.")
    rhomat = np.zeros((n_pop, n_pop))
    for i in range(n_pop - 1):
        for j in range(i + 1, n_pop):
            ind = np.where((pr["popd"] == up[i]) | (pr["popd"] == up[j]))[0]
            fit = fitGP(y=pr["yds"][ind], x=pr["xds"][ind], pop=pr["popd"][ind], time=pr["timed"][ind],
                        scaling="none", initpars=None if initpars is None else np.array(initpars, dtype=np.float64),
                        modeprior=modeprior, fixedpars=fixedpars)
            rhomat[i, j] = fit["pars"][pr["d"] + 2]
    return rhomat + rhomat.T + np.eye(n_pop), list(up)
