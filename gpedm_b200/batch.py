"""Batched independent fits (E/tau model-selection grids, multistart, per-population models):
the path that shards across GPUs.  Mirrors the user-level loops of the reference
(vignettes/choosingE.Rmd:63-86, R/rhomatrix.R:169-182) as one call into `gpedm_fit_batch`.
"""
import ctypes as C

import numpy as np

from ._lib import FitDesc, Result, check, dptr, f64, iptr, lib
from .core import SIGMA2MAX, SIGMA2MIN, VEMAX, VEMIN, logit, pop_codes
from .fitgp import _lagmatrix, _mean, _sd, getR2


def default_initparst(d, single_pop=True):
    """Default start of fitGP (R/GPfunctions.R:409, :425-429)."""
    initpars = np.concatenate([np.full(d, 0.1), [0.001, 1 - 0.001, 0.5]])
    t = np.concatenate([np.log(initpars[:d]), [logit(initpars[d], VEMIN, VEMAX),
                                               logit(initpars[d + 1], SIGMA2MIN, SIGMA2MAX),
                                               logit(initpars[d + 2], 0.0, 1.0)]])
    if single_pop:
        t[d + 2] = 0.0
    return t


def fit_batch(problems, ngpus=1, devices=None, want_loo=False):
    """problems: list of dicts with Y (N), X (N x d), optional pop, rhomatrix, initparst, fixedpars,
    modeprior, rhofixed.  Returns a list of result dicts (plus loo_mean / loo_var / insamp_mean when
    want_loo)."""
    n = len(problems)
    descs = (FitDesc * n)()
    results = (Result * n)()
    keep = []
    outs = []
    for i, pr in enumerate(problems):
        Y = f64(pr["Y"]).ravel()
        X = np.asarray(pr["X"], dtype=np.float64)
        if X.ndim == 1:
            X = X[:, None]
        X = f64(X)
        N, d = X.shape
        pop = pr.get("pop")
        codes, levels = pop_codes(np.ones(N, dtype=np.int64) if pop is None else pop)
        rm = pr.get("rhomatrix")
        rm = None if rm is None else f64(rm)
        ip = pr.get("initparst")
        ip = f64(default_initparst(d, len(levels) == 1) if ip is None else ip).ravel()
        fp = pr.get("fixedpars")
        fp = None if fp is None else f64(fp).ravel()
        o = dict(loo_mean=np.empty(N), loo_var=np.empty(N), insamp_mean=np.empty(N), insamp_var=np.empty(N))
        keep.append((Y, X, codes, rm, ip, fp, o))
        ds = descs[i]
        ds.N, ds.d, ds.np = N, d, len(levels)
        ds.X, ds.Y, ds.pop = dptr(X), dptr(Y), iptr(codes)
        ds.rhomat, ds.initparst, ds.fixedpars = dptr(rm), dptr(ip), dptr(fp)
        ds.modeprior = float(pr.get("modeprior", 1.0))
        rf = pr.get("rhofixed")
        ds.rhofixed = float("nan") if rf is None else float(rf)
        ds.want_loo = 1 if want_loo else 0
        if want_loo:
            ds.loo_mean, ds.loo_var = dptr(o["loo_mean"]), dptr(o["loo_var"])
            ds.insamp_mean, ds.insamp_var = dptr(o["insamp_mean"]), dptr(o["insamp_var"])
        outs.append(o)
    dev = None if devices is None else np.ascontiguousarray(np.asarray(devices, dtype=np.int32))
    check(lib().gpedm_fit_batch(n, descs, int(ngpus), iptr(dev), None, results))
    res = []
    for i in range(n):
        r = results[i].to_dict()
        if want_loo:
            r.update(outs[i])
        res.append(r)
    return res


def grid_problems(y, Es, taus, pop=None, scaling="global"):
    """Training sets of the E x tau model-selection grid of vignettes/choosingE.Rmd:63-86:
    for each (E, tau) lag y, scale (global), drop incomplete rows.  Returns (problems, meta)."""
    y = np.asarray(y, dtype=np.float64)
    popv = np.ones(len(y), dtype=np.int64).astype(object) if pop is None else np.asarray(pop, dtype=object)
    if scaling != "global":
        raise NotImplementedError("grid helper supports global scaling")
    problems, meta = [], []
    for E in Es:
        for tau in taus:
            x = _lagmatrix(y, popv, E, tau)
            ym, ys = _mean(y), _sd(y)
            yds = (y - ym) / ys
            xm = np.array([_mean(x[:, j]) for j in range(x.shape[1])])
            xs = np.array([_sd(x[:, j]) for j in range(x.shape[1])])
            xds = (x - xm[None, :]) / xs[None, :]
            ok = ~(np.isnan(yds) | np.isnan(xds).any(axis=1))
            problems.append(dict(Y=yds[ok], X=xds[ok], pop=popv[ok]))
            meta.append(dict(E=E, tau=tau, ymean=ym, ysd=ys, rows=ok, yobs=y[ok]))
    return problems, meta


def fit_grid(y, Es, taus, pop=None, ngpus=1, devices=None, scaling="global", modeprior=1.0):
    """fitGP(y, E, tau, predictmethod="loo") over an E x tau grid, returning one dict per cell with
    the fitted modes and the leave-one-out R2 (the model-selection criterion of choosingE.Rmd).
    One call into `gpedm_fit_grid`: the raw series is uploaded once per GPU and every cell's lags,
    scaling and complete-row selection are built on the device."""
    from ._lib import SCALE_CODES, FitStats
    y = f64(y).ravel()
    T = len(y)
    cells = [(int(E), int(tau)) for E in Es for tau in taus]
    n = len(cells)
    Ea = np.ascontiguousarray([c[0] for c in cells], dtype=np.int32)
    ta = np.ascontiguousarray([c[1] for c in cells], dtype=np.int32)
    codes, levels = pop_codes(np.ones(T, dtype=np.int64) if pop is None else pop)
    results = (Result * n)()
    stats = (FitStats * n)()
    dev = None if devices is None else np.ascontiguousarray(np.asarray(devices, dtype=np.int32))
    check(lib().gpedm_fit_grid(T, dptr(y), iptr(codes), len(levels), n, iptr(Ea), iptr(ta), SCALE_CODES[scaling],
                               float(modeprior), int(ngpus), iptr(dev), None, results, stats))
    out = []
    for (E, tau), r, st in zip(cells, results, stats):
        r = r.to_dict()
        out.append(dict(E=E, tau=tau, pars=r["pars"], iter=r["iter"], nevals=r["nevals"], ln_post=-r["nllpost"],
                        lnL_LOO=r["lnL_LOO"], df=r["df"], R2_insample=st.R2_insample, R2_loo=st.R2_loo,
                        rmse_insample=st.rmse_insample, rmse_loo=st.rmse_loo, N=int(st.n), status=r["status"]))
    return out


def fit_grid_hostprep(y, Es, taus, pop=None, ngpus=1, devices=None):
    """Same grid with the training sets prepared on the host (`grid_problems`) and fitted through
    `gpedm_fit_batch` - the cross-check of the on-device preparation."""
    problems, meta = grid_problems(y, Es, taus, pop)
    res = fit_batch(problems, ngpus=ngpus, devices=devices, want_loo=True)
    out = []
    for r, m in zip(res, meta):
        loo = r["loo_mean"] * m["ysd"] + m["ymean"]
        ins = r["insamp_mean"] * m["ysd"] + m["ymean"]
        out.append(dict(E=m["E"], tau=m["tau"], pars=r["pars"], iter=r["iter"], nevals=r["nevals"],
                        ln_post=-r["nllpost"], R2_insample=getR2(m["yobs"], ins), R2_loo=getR2(m["yobs"], loo),
                        status=r["status"]))
    return out


def initpars_to_parst(initpars, d, single_pop):
    """fitGP's transform of a constrained start (phi_1..d, ve, sigma2, rho) to the optimiser's scale
    (R/GPfunctions.R:421-429; note rhomin = 0 here, 1e-4 inside getlikegrad)."""
    ip = np.asarray(initpars, dtype=np.float64)
    if len(ip) != d + 3:
        raise ValueError("initpars must have number of predictors + 3 entries")
    with np.errstate(divide="ignore"):
        t = np.concatenate([np.log(ip[:d]), [logit(ip[d], VEMIN, VEMAX), logit(ip[d + 1], SIGMA2MIN, SIGMA2MAX),
                                             logit(ip[d + 2], 0.0, 1.0)]])
    if single_pop:
        t[d + 2] = 0.0
    return t


def multistart(Y, X, pop=None, starts=None, nstarts=8, seed=0, ngpus=1, devices=None, **fit_kw):
    """The same training set fitted from several starting points in ONE batched call - the remedy the
    reference suggests against local minima ("One could try using different values of initpars",
    vignettes/hierarchical.Rmd:47), which in R is a loop of fitGP(..., initpars=) calls.
    starts: list of constrained initpars vectors (phi_1..d, ve, sigma2, rho); default: fitGP's own start
    (R/GPfunctions.R:409) plus nstarts-1 random ones (phi log-uniform in [0.01, 2], ve in [0.001, 0.5],
    sigma2 in [0.5, 4], rho in [0.1, 0.9]).  Returns (results sorted by ln_post descending, index of best)."""
    X = np.asarray(X, dtype=np.float64)
    if X.ndim == 1:
        X = X[:, None]
    d = X.shape[1]
    single = pop is None or len(set(np.asarray(pop, dtype=object).tolist())) == 1
    if starts is None:
        rng = np.random.default_rng(seed)
        starts = [np.concatenate([np.full(d, 0.1), [0.001, 1 - 0.001, 0.5]])]
        for _ in range(nstarts - 1):
            starts.append(np.concatenate([np.exp(rng.uniform(np.log(0.01), np.log(2.0), d)), [rng.uniform(0.001, 0.5)],
                                          [rng.uniform(0.5, 4.0)], [rng.uniform(0.1, 0.9)]]))
    probs = [dict(Y=Y, X=X, pop=pop, initparst=initpars_to_parst(s, d, single), **fit_kw) for s in starts]
    res = fit_batch(probs, ngpus=ngpus, devices=devices)
    for r, s in zip(res, starts):
        r["initpars"] = np.asarray(s, dtype=np.float64)
        r["ln_post"] = -r["nllpost"]
    order = sorted(range(len(res)), key=lambda i: -res[i]["ln_post"] if res[i]["status"] == 0 else np.inf)
    return [res[i] for i in order], order[0]


def b_profile_problems(y, m, h, bs, z=None, pop=None):
    """Training sets of the fisheries model's catchability search: for every candidate b the composite
    predictor is x = m - b h (escapement, R/GPfisheries.R:247-251), then fitGP's global scaling and
    complete-row selection (R/GPfunctions.R:378-385, :445-451); ytrans = "none".  The reference evaluates
    these one b at a time inside optimize() (R/GPfisheries.R:114-117, objective GPnlike_fish :232-294)."""
    y = np.asarray(y, dtype=np.float64)
    m = np.asarray(m, dtype=np.float64).reshape(len(y), -1)
    h = np.asarray(h, dtype=np.float64).reshape(len(y), -1)
    if m.shape != h.shape:
        raise ValueError("m and h must have the same number of columns")
    zc = None if z is None else np.asarray(z, dtype=np.float64).reshape(len(y), -1)
    popv = np.ones(len(y), dtype=np.int64).astype(object) if pop is None else np.asarray(pop, dtype=object)
    probs = []
    for b in bs:
        x = m - float(b) * h
        if zc is not None:
            x = np.column_stack([x, zc])
        yds = (y - _mean(y)) / _sd(y)
        xds = (x - np.array([_mean(x[:, j]) for j in range(x.shape[1])])[None, :]) / \
            np.array([_sd(x[:, j]) for j in range(x.shape[1])])[None, :]
        ok = ~(np.isnan(yds) | np.isnan(xds).any(axis=1))
        probs.append(dict(Y=yds[ok], X=xds[ok], pop=popv[ok]))
    return probs


def b_profile(y, m, h, bs, z=None, pop=None, ngpus=1, devices=None, **fit_kw):
    """ln_post of the fisheries model over a grid of catchabilities b, all fits in ONE batched call; the
    maximiser is what fitGP_fish's optimize() searches for (R/GPfisheries.R:114-117).  Returns a list of
    dicts (b, ln_post, pars, iter, status) and the index of the best b."""
    probs = b_profile_problems(y, m, h, bs, z, pop)
    for p in probs:
        p.update(fit_kw)
    res = fit_batch(probs, ngpus=ngpus, devices=devices)
    out = [dict(b=float(b), ln_post=-r["nllpost"], pars=r["pars"], iter=r["iter"], nevals=r["nevals"], status=r["status"])
           for b, r in zip(bs, res)]
    best = max(range(len(out)), key=lambda i: out[i]["ln_post"] if out[i]["status"] == 0 else -np.inf)
    return out, best
