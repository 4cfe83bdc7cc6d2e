"""`fitGP` / `predict` / `update` / `predict_seq` / `makelags`: the reference's user API
(R/GPfunctions.R:221-520, :880-1341, :1526-1857; R/predict_seq.R:38-99) as a Python host layer
over the CUDA library.  Data preparation (lags, scaling, NA removal, back-transforms, fit
statistics) is the O(N d) part the reference keeps in R and stays on the host here; every
N^2 / N^3 step goes to the GPU through `core.GPModel`.

A "data frame" is any mapping column-name -> 1-D array (dict, pandas.DataFrame, ...).
Not mirrored (out of the hot path): `linprior` (an `lm()` pre-processing of y),
the fisheries wrapper, `vtimestep` / `augment` lag generation, plotting.
"""
import math

import numpy as np

from . import _lib
from .core import GPModel, SIGMA2MAX, SIGMA2MIN, VEMAX, VEMIN, logit, unique_stable


# ------------------------------------------------------------------------------ small helpers
def _mean(v):
    v = np.asarray(v, dtype=np.float64)
    v = v[~np.isnan(v)]
    return float(np.mean(v)) if v.size else float("nan")


def _sd(v):
    v = np.asarray(v, dtype=np.float64)
    v = v[~np.isnan(v)]
    return float(np.std(v, ddof=1)) if v.size > 1 else float("nan")


def getR2(obs, pred):
    """R/GPfunctions.R:1366-1370."""
    obs = np.asarray(obs, dtype=np.float64)
    pred = np.asarray(pred, dtype=np.float64)
    ok = ~(np.isnan(obs) | np.isnan(pred))
    o, p = obs[ok], pred[ok]
    with np.errstate(divide="ignore", invalid="ignore"):
        return float(1.0 - np.sum((o - p) ** 2) / np.sum((o - np.mean(o)) ** 2))


def getR2pop(obs, pred, pop, type="within"):
    """R/GPfunctions.R:1394-1437."""
    if type not in ("within", "centered", "scaled"):
        raise ValueError("type must be one of within, centered, scaled")
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


def _names(x):
    return [x] if isinstance(x, str) else list(x)


def _as_matrix(data, keys):
    return np.column_stack([np.asarray(data[k], dtype=np.float64) for k in _names(keys)])


def _rbind(a, b):
    return {k: np.concatenate([np.asarray(a[k]), np.asarray(b[k])]) for k in a}


def _rows(df, mask):
    return {k: np.asarray(df[k])[mask] for k in df}


# ------------------------------------------------------------------------------ makelags
def makelags(data=None, y=None, pop=None, E=None, tau=None, yname=None, forecast=False, time=None,
             append=False):
    """Standard (fixed-timestep) delay vectors, R/GPfunctions.R:1526-1645 + makelags_subrt :1824-1857.
    Returns a dict of named columns `<var>_<lag>` (plus pop / time columns when forecast=True,
    :1622-1643)."""
    if E is None or tau is None or E < 0 or tau < 0 or int(E) != E or int(tau) != tau:
        raise ValueError("E and tau must be positive integers")
    E, tau = int(E), int(tau)
    pname, tname = "pop", "time"
    if data is not None:
        ynames = _names(y)
        yv = _as_matrix(data, y)
        if pop is not None:
            pname = pop
            pop = np.asarray(data[pop], dtype=object)
        if time is not None:
            tname = time
            time = np.asarray(data[time], dtype=np.float64)
    else:
        yv = np.asarray(y, dtype=np.float64)
        if yv.ndim == 1:
            yv = yv[:, None]
        nv = yv.shape[1]
        if yname is None:
            yname = "var"
        ynames = _names(yname)
        if len(ynames) != nv:
            ynames = ["%s%d" % (ynames[0], i + 1) for i in range(nv)]
    n, nv = yv.shape
    if pop is None:
        pop = np.ones(n, dtype=np.int64)
    pop = np.asarray(pop, dtype=object)
    up = unique_stable(pop)
    out = np.full((n, nv * E), np.nan)
    outfore = np.full((len(up) * tau, nv * E), np.nan)
    popfore = np.repeat(np.array(up, dtype=object), tau)
    colnames = [None] * (nv * E)
    for u in up:
        col = 0
        ind = np.where(pop == u)[0]
        indfore = np.where(popfore == u)[0]
        for j in range(nv):
            ts = yv[ind, j]
            if len(ts) <= E * tau:
                raise ValueError("time series length <= E*tau for pop %s" % u)
            for i in range(1, E + 1):
                out[ind, col] = np.concatenate([np.full(tau * i, np.nan), ts[: len(ts) - tau * i]])
                outfore[indfore, col] = ts[len(ts) - tau * i: len(ts) - tau * (i - 1)]
                colnames[col] = "%s_%d" % (ynames[j], i * tau)
                col += 1
    if not forecast:
        res = {c: out[:, k] for k, c in enumerate(colnames)}
        if append and data is not None:
            merged = {k: np.asarray(data[k]) for k in data}
            merged.update(res)
            return merged
        return res
    res = {}
    if time is not None:
        timefore = np.zeros(len(up) * tau)
        for u in up:
            ti = time[pop == u]
            tdiff = ti[-1] - ti[-2]
            timefore[popfore == u] = ti[-1] + np.arange(1, tau + 1) * tdiff
        res[tname] = timefore
    if len(up) > 1:
        res[pname] = popfore
    res.update({c: outfore[:, k] for k, c in enumerate(colnames)})
    return res


def _lagmatrix(xv, pop, E, tau):
    d = makelags(y=xv, pop=pop, E=E, tau=tau)
    return np.column_stack(list(d.values()))


# ------------------------------------------------------------------------------ GP object
class GP(dict):
    """Fitted model: a dict with the reference's list entries (R/GPfunctions.R:490-519):
    pars, parst, grad, iter, inputs, scaling, insampresults, insampfitstats, [insampfitstatspop],
    [outsampresults, outsampfitstats, outsampfitstatspop, GPgrad], call.  `covm` (the four dense
    N x N matrices the reference stores, :638-645, :815) is materialised from the device on
    first access only."""

    def __getitem__(self, key):
        if key == "covm" and not dict.__contains__(self, "covm"):
            m = dict.__getitem__(self, "model")
            self["covm"] = dict(popsame=m.matrix(_lib.MAT_POPSAME), Cd=m.matrix(_lib.MAT_CD),
                                Sigma=m.matrix(_lib.MAT_SIGMA), iKVs=m.matrix(_lib.MAT_IKVS))
        return dict.__getitem__(self, key)

    def close(self):
        m = dict.get(self, "model")
        if m is not None:
            m.close()


def fitGP(data=None, y=None, x=None, pop=None, time=None, E=None, tau=None, scaling="global", initpars=None,
          modeprior=1, fixedpars=None, rhofixed=None, rhomatrix=None, rhomatrix_names=None, augdata=None,
          linprior="none", predictmethod=None, newdata=None, xnew=None, popnew=None, timenew=None, ynew=None,
          returnGPgrad=False, exclradius=0, device=0, _prep_only=False):
    """Fit a GP model (reference fitGP, R/GPfunctions.R:221-520)."""
    call = dict(data=data, y=y, x=x, pop=pop, time=time, E=E, tau=tau, scaling=scaling, initpars=initpars,
                modeprior=modeprior, fixedpars=fixedpars, rhofixed=rhofixed, rhomatrix=rhomatrix,
                rhomatrix_names=rhomatrix_names, augdata=augdata, linprior=linprior, predictmethod=predictmethod,
                newdata=newdata, xnew=xnew, popnew=popnew, timenew=timenew, ynew=ynew, returnGPgrad=returnGPgrad,
                exclradius=exclradius, device=device)
    if (E is None) != (tau is None):
        raise ValueError("Both E and tau must be supplied if generating lags internally.")
    if x is None and (E is None or tau is None):
        raise ValueError("x and/or E and tau must be supplied")
    if rhofixed is not None:
        if rhomatrix is not None:
            raise ValueError("Supply either rhofixed or rhomatrix, not both")
        rhofixed = min(max(float(rhofixed), 0.0001), 0.9999)  # :241-248
    if scaling not in ("global", "local", "none"):
        raise ValueError("scaling must be one of global, local, none")
    if linprior != "none":
        raise NotImplementedError("linprior is host-side lm() pre-processing in the reference and is not mirrored")
    inputs = {}
    if data is not None:  # :261-276
        inputs["y_names"] = y
        yv = np.asarray(data[y], dtype=np.float64)
        xv = None
        if x is not None:
            inputs["x_names"] = _names(x)
            xv = _as_matrix(data, x)
        popv = None
        if pop is not None:
            inputs["pop_names"] = pop
            popv = np.asarray(data[pop], dtype=object)
        timev = None
        if time is not None:
            inputs["time_names"] = time
            timev = np.asarray(data[time], dtype=np.float64)
    else:
        yv = np.asarray(y, dtype=np.float64)
        xv = None if x is None else np.asarray(x, dtype=np.float64)
        popv = None if pop is None else np.asarray(pop, dtype=object)
        timev = None if time is None else np.asarray(time, dtype=np.float64)
    n = len(yv)
    if popv is None:
        popv = np.ones(n, dtype=np.int64).astype(object)
    if timev is None:  # :283-289
        timev = np.zeros(n)
        for u in unique_stable(popv):
            m = popv == u
            timev[m] = np.arange(1, m.sum() + 1)
    if E is not None:  # :292-306
        if xv is None:
            xv = yv
        xv = _lagmatrix(xv, popv, E, tau)
        inputs["E"], inputs["tau"] = E, tau
    if xv.ndim == 1:
        xv = xv[:, None]
    d = xv.shape[1]
    if augdata is not None:  # :313-330
        xaug = _as_matrix(augdata, inputs["x_names"])
        yaug = np.asarray(augdata[inputs["y_names"]], dtype=np.float64)
        timeaug = np.asarray(augdata[inputs["time_names"]], dtype=np.float64)
        popaug = (np.ones(len(yaug), dtype=np.int64).astype(object) if "pop_names" not in inputs
                  else np.asarray(augdata[inputs["pop_names"]], dtype=object))
        yd = np.concatenate([yv, yaug])
        xd = np.vstack([xv, xaug])
        timed = np.concatenate([timev, timeaug])
        popd = np.concatenate([popv, popaug])
        primary = np.concatenate([np.ones(n, bool), np.zeros(len(yaug), bool)])
    else:
        yd, xd, timed, popd, primary = yv, xv, timev, popv, np.ones(n, bool)
    up = unique_stable(popd)
    # rescale data :359-405
    if scaling in ("none", "global"):
        ymeans, ysds = _mean(yd), _sd(yd)
        xmeans = np.array([_mean(xd[:, j]) for j in range(d)])
        xsds = np.array([_sd(xd[:, j]) for j in range(d)])
        if scaling == "global":
            yds = (yd - ymeans) / ysds
            xds = (xd - xmeans[None, :]) / xsds[None, :]
        else:
            yds, xds = yd.copy(), xd.copy()
    else:
        ymeans = {u: _mean(yd[popd == u]) for u in up}
        ysds = {u: _sd(yd[popd == u]) for u in up}
        xmeans = {u: np.array([_mean(xd[popd == u, j]) for j in range(d)]) for u in up}
        xsds = {u: np.array([_sd(xd[popd == u, j]) for j in range(d)]) for u in up}
        yds, xds = yd.copy(), xd.copy()
        for u in up:
            m = popd == u
            yds[m] = (yd[m] - ymeans[u]) / ysds[u]
            xds[m] = (xd[m] - xmeans[u][None, :]) / xsds[u][None, :]
    if _prep_only:  # the section getrhomatrix duplicates from fitGP (R/rhomatrix.R:26-160)
        return dict(yds=yds, xds=xds, popd=popd, timed=timed, primary=primary, up=up, d=d)
    if initpars is None:  # :408-410
        initpars = np.concatenate([np.full(d, 0.1), [0.001, 1 - 0.001, 0.5]])
    initpars = np.array(initpars, dtype=np.float64)
    if len(initpars) != d + 3:
        raise ValueError("initpars must have number of predictors + 3 entries")
    fp = np.full(d + 2, np.nan) if fixedpars is None else np.asarray(fixedpars, dtype=np.float64)
    if len(fp) != d + 2:
        raise ValueError("Length of fixedpars must be number of predictors + 2 (phi, ve, sigma2)")
    fi = ~np.isnan(fp)
    initpars[:d + 2][fi] = fp[fi]
    with np.errstate(divide="ignore"):
        initparst = np.concatenate([np.log(initpars[:d]), [logit(initpars[d], VEMIN, VEMAX),  # :425-426
                                                           logit(initpars[d + 1], SIGMA2MIN, SIGMA2MAX),
                                                           logit(initpars[d + 2], 0.0, 1.0)]])
    if len(up) == 1:
        initparst[d + 2] = 0.0  # :429
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
    # optimise model on the GPU :454
    model = GPModel(Y, X, Pop, rhomatrix, rhomatrix_names, device)
    try:
        res = model.fit_rprop(initparst, modeprior, None if not fi.any() else fp, rhofixed)
        mean = model.vector(_lib.VEC_MEAN)
        covdiag = model.vector(_lib.VEC_VAR)
    except Exception:
        model.close()
        raise
    pars = res["pars"]
    ve = pars[d]
    ypred = np.full(len(yd), np.nan)
    yfsd = np.full(len(yd), np.nan)
    ysd = np.full(len(yd), np.nan)
    ypred[completerows] = mean  # :458-461
    with np.errstate(invalid="ignore"):
        yfsd[completerows] = np.sqrt(covdiag)
        ysd[completerows] = np.sqrt(covdiag + ve)
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
    out = GP()
    names = ["phi%d" % (i + 1) for i in range(d)] + ["ve", "sigma2", "rho"]
    out["par_names"] = names
    out["pars"] = pars
    out["parst"] = res["parst"]
    out["grad"] = res["grad"]
    out["iter"] = res["iter"]
    out["nevals"] = res["nevals"]
    out["model"] = model
    inputs.update(dict(y=yv, x=xv, yd=yd, xd=xd, pop=popv, time=timev, completerows=completerows,
                       Y=Y, X=X, Pop=Pop, Time=Time, Primary=Primary))
    out["inputs"] = inputs
    out["scaling"] = dict(scaling=scaling, ymeans=ymeans, ysds=ysds, xmeans=xmeans, xsds=xsds)
    out["insampresults"] = dict(timestep=timev, pop=popv, predmean=ypred, predfsd=yfsd, predsd=ysd, obs=yv)
    stats = dict(R2=getR2(yv, ypred), rmse=math.sqrt(_mean((yv - ypred) ** 2)), ln_post=-res["nllpost"],
                 ln_prior=res["lp"], lnL_LOO=res["lnL_LOO"], df=res["df"], SS=res["SS"], logdet=res["logdet"])
    if len(unique_stable(popv)) > 1:  # :502-507
        out["insampfitstatspop"] = getR2pop(yv, ypred, popv)
        stats["R2centered"] = getR2pop(yv, ypred, popv, "centered")
        stats["R2scaled"] = getR2pop(yv, ypred, popv, "scaled")
    out["insampfitstats"] = stats
    if predictmethod is not None or xnew is not None or newdata is not None:  # :512-515
        out.update(predict(out, predictmethod, newdata, xnew, popnew, timenew, ynew, returnGPgrad, exclradius))
    out["call"] = call
    return out


def predict(obj, predictmethod="loo", newdata=None, xnew=None, popnew=None, timenew=None, ynew=None,
            returnGPgrad=False, exclradius=0):
    """Predictions from a fitted model (reference predict.GP, R/GPfunctions.R:880-1341)."""
    model = obj["model"]
    pars = obj["pars"]
    inp, sc = obj["inputs"], obj["scaling"]
    yv = inp["y"]
    d = inp["X"].shape[1]
    ve = pars[d]
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
        if ynew is not None:
            Tslp = len(ynew)
        else:
            xa = np.asarray(xnew)
            Tslp = xa.shape[0]
        if popnew is None:
            popnew = np.ones(Tslp, dtype=np.int64).astype(object)
        popnew = np.asarray(popnew, dtype=object)
        if timenew is None:
            timenew = np.zeros(Tslp)
            for u in unique_stable(popnew):
                m = popnew == u
                timenew[m] = np.arange(1, m.sum() + 1)
        if inp.get("E") is not None:  # :966-971
            if xnew is None:
                xnew = ynew
            xnew = _lagmatrix(np.asarray(xnew, dtype=np.float64), popnew, inp["E"], inp["tau"])
        xnew = np.asarray(xnew, dtype=np.float64)
        if xnew.ndim == 1:
            xnew = xnew[:, None]
        xnews = xnew.copy()
        if scaling == "global":
            xnews = (xnew - xmeans[None, :]) / xsds[None, :]
        if scaling == "local":
            for u in unique_stable(popnew):
                m = popnew == u
                if u not in xmeans:
                    raise ValueError("population %r not present in the training data" % (u,))
                xnews[m] = (xnew[m] - xmeans[u][None, :]) / xsds[u][None, :]
        comp = ~np.isnan(xnews).any(axis=1)  # :999
        ypred = np.full(Tslp, np.nan)
        yfsd = np.full(Tslp, np.nan)
        ysd = np.full(Tslp, np.nan)
        if comp.any():
            ymean, yvar, G = model.predict_new(xnews[comp], popnew[comp], returnGPgrad)
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
        if predictmethod not in ("loo", "lto", "sequential"):
            raise ValueError("predictmethod must be one of loo, lto, sequential")
        Time, Primary = inp["Time"], inp["Primary"]
        cr = inp["completerows"][:len(yv)]
        ypred = np.full(len(yv), np.nan)
        yfsd = np.full(len(yv), np.nan)
        ysd = np.full(len(yv), np.nan)
        if predictmethod in ("loo", "lto"):
            if predictmethod == "loo":
                ymean, yvar, G = model.predict_loo(exclradius, Time, Primary, returnGPgrad)
            else:
                ymean, yvar, G = model.predict_lto(Time, exclradius, Primary, returnGPgrad)
            ypred[cr] = ymean
            with np.errstate(invalid="ignore"):
                yfsd[cr] = np.sqrt(yvar)
                ysd[cr] = np.sqrt(yvar + ve)
            if returnGPgrad:
                gpgrad = np.full((len(yv), d), np.nan)
                gpgrad[cr] = G
        else:
            ymean, ysd2 = model.predict_sequential(Time)  # :1196-1251
            ypred[cr] = ymean[Primary]
            with np.errstate(invalid="ignore"):
                yfsd[cr] = np.sqrt(ysd2 ** 2 - ve)[Primary]
            ysd[cr] = ysd2[Primary]
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


def update(obj, **changes):
    """Re-run the stored call with some arguments replaced (R `update()` on the stored call,
    R/GPfunctions.R:228, :517)."""
    call = dict(obj["call"])
    call.update(changes)
    return fitGP(**call)


def predict_seq(obj, newdata, restart_initpars=False):
    """R/predict_seq.R:38-99: for each new time value predict, append the observations, and refit
    with `update()` warm-started from the previous posterior mode."""
    inp = obj["inputs"]
    if inp.get("time_names") is None:
        raise ValueError("Model must include `data` and `time` to use this function.")
    if inp.get("E") is not None:
        raise ValueError("Lags must be pre-generated to use this function. See option A1 in help(fitGP).")
    tn = inp["time_names"]
    newtimes = unique_stable(np.asarray(newdata[tn]).tolist())
    model = obj
    dataup = {k: np.asarray(obj["call"]["data"][k]) for k in obj["call"]["data"]}
    preds = []
    for t in newtimes:
        nd = _rows(newdata, np.asarray(newdata[tn]) == t)
        preds.append(predict(model, newdata=nd)["outsampresults"])
        dataup = _rbind(dataup, {k: nd[k] for k in dataup})
        inits = None if restart_initpars else model["pars"]
        new = update(model, data=dataup, initpars=inits, newdata=None, predictmethod=None)
        if model is not obj:
            model.close()
        model = new
    res = preds[0]
    for p in preds[1:]:
        res = _rbind(res, p)
    model["outsampresults"] = res
    model["outsampfitstats"] = dict(R2=getR2(res["obs"], res["predmean"]),
                                    rmse=math.sqrt(_mean((res["obs"] - res["predmean"]) ** 2)))
    if len(unique_stable(res["pop"])) > 1:
        model["outsampfitstatspop"] = getR2pop(res["obs"], res["predmean"], res["pop"])
        model["outsampfitstats"]["R2centered"] = getR2pop(res["obs"], res["predmean"], res["pop"], "centered")
        model["outsampfitstats"]["R2scaled"] = getR2pop(res["obs"], res["predmean"], res["pop"], "scaled")
    return model


def predict_iter(obj, newdata, xlags=None):
    """Iterated multi-step forecasts (reference R/predict_iter.R:57-163, non-fisheries models): predict one
    time step, feed the prediction in as the first lag of the next step, shift the other lags."""
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
    has_pop = inp.get("pop_names") is not None
    pn = inp["pop_names"] if has_pop else "pop"
    if not has_pop:
        nd[pn] = np.ones(len(nd[tn]), dtype=np.int64).astype(object)
    elif pn not in nd:
        raise ValueError("Pop column '%s' not found in newdata." % pn)
    newtimes = unique_stable(nd[tn].tolist())
    up = unique_stable(nd[pn])
    preds = []
    for i, t in enumerate(newtimes):
        mask = nd[tn] == t
        cur = _rows(nd, mask)
        if not has_pop:
            cur = {k: v for k, v in cur.items() if k != pn}
        pr = predict(obj, newdata=cur)["outsampresults"]
        preds.append(pr)
        if np.any(np.isinf(pr["predmean"])):
            break
        if i + 1 < len(newtimes):
            nxt = nd[tn] == newtimes[i + 1]
            for u in up:
                pj = pr["predmean"][np.asarray(pr["pop"], dtype=object) == u] if has_pop else pr["predmean"]
                sel_n, sel_c = nxt & (nd[pn] == u), mask & (nd[pn] == u)
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
        if len(unique_stable(res["pop"])) > 1:
            out["outsampfitstatspop"] = getR2pop(res["obs"], res["predmean"], res["pop"])
    return out


def getconditionals(fit, xrange="default", extrap=0.01, nvals=25):
    """Conditional responses (reference R/PlotsConditionals.R:166-266, without the plot): for every
    population and predictor, `nvals` points along that predictor with the other scaled predictors at
    0.  All np x d x nvals slice points go to the GPU in ONE predict_new call against the resident
    factorisation (the reference issues np*d getcov calls and np*d*nvals quadratic forms).
    Returns one dict per population: poppred, xval, predmean, predsd (each nvals x d)."""
    inp, sc = fit["inputs"], fit["scaling"]
    X, Pop = inp["X"], np.asarray(inp["Pop"], dtype=object)
    d = X.shape[1]
    scaling = sc["scaling"]
    if xrange not in ("default", "local", "global"):
        raise ValueError("xrange must be one of default, local, global")
    xr = ("local" if scaling == "local" else "global") if xrange == "default" else xrange
    up = unique_stable(Pop)
    pts, pops, grids = [], [], {}
    for u in up:
        indi = Pop == u
        for i in range(d):
            if xr == "local":
                ext = (X[indi, i].max() - X[indi, i].min()) * extrap
                xgi = np.linspace(X[indi, i].min() - ext, X[indi, i].max() + ext, nvals)
            else:  # quirk kept: upper end from max over the whole X matrix (:205)
                ext = (X[:, i].max() - X[:, i].min()) * extrap
                xgi = np.linspace(X[:, i].min() - ext, X.max() + ext, nvals)
            xp = np.zeros((nvals, d))
            xp[:, i] = xgi
            pts.append(xp)
            pops.extend([u] * nvals)
            grids[(u, i)] = xgi
    mean, var, _ = fit["model"].predict_new(np.vstack(pts), np.array(pops, dtype=object))
    out, k = [], 0
    for u in up:
        xval, pm, ps = np.zeros((nvals, d)), np.zeros((nvals, d)), np.zeros((nvals, d))
        for i in range(d):
            pm[:, i] = mean[k:k + nvals]
            with np.errstate(invalid="ignore"):
                ps[:, i] = np.sqrt(var[k:k + nvals])
            xval[:, i] = grids[(u, i)]
            k += nvals
        if scaling == "global":
            pm = pm * sc["ysds"] + sc["ymeans"]
            ps = np.sqrt(ps ** 2 * sc["ysds"] ** 2)
            xval = xval * sc["xsds"][None, :] + sc["xmeans"][None, :]
        if scaling == "local":
            pm = pm * sc["ysds"][u] + sc["ymeans"][u]
            ps = np.sqrt(ps ** 2 * sc["ysds"][u] ** 2)
            xval = xval * sc["xsds"][u][None, :] + sc["xmeans"][u][None, :]
        out.append(dict(poppred=np.array([u] * nvals, dtype=object), xval=xval, predmean=pm, predsd=ps))
    return out


def summary(obj):
    """Text summary in the layout of summary.GP (R/PlotsConditionals.R:10-52)."""
    d = obj["inputs"]["X"].shape[1]
    pars = obj["pars"]
    lines = ["Number of predictors: %d" % d, "Length scale parameters:"]
    for i in range(d):
        lines.append("  phi%d %.5f" % (i + 1, pars[i]))
    lines.append("Process variance (ve): %.7g" % pars[d])
    lines.append("Pointwise prior variance (sigma2): %.7g" % pars[d + 1])
    npop = len(unique_stable(obj["inputs"]["pop"]))
    lines.append("Number of populations: %d" % npop)
    if npop > 1:
        lines.append("Dynamic correlation (rho): %.7g" % pars[d + 2])
    lines.append("In-sample R-squared: %.7g" % obj["insampfitstats"]["R2"])
    if "outsampfitstats" in obj:
        lines.append("Out-of-sample R-squared: %.7g" % obj["outsampfitstats"]["R2"])
    return "\n".join(lines)


def getrhomatrix(data=None, y=None, x=None, pop=None, time=None, E=None, tau=None, scaling="global",
                 initpars=None, modeprior=1, fixedpars=None, augdata=None, ngpus=1, devices=None):
    """Pairwise dynamic-correlation matrix (reference getrhomatrix, R/rhomatrix.R:22-186): the data
    preparation of fitGP, then a two-population hierarchical fit for every pair of populations
    (`:169-182`, scaling "none" on the already scaled data) whose rho fills a symmetric matrix with unit
    diagonal.  The np(np-1)/2 pair fits are independent: ONE `gpedm_fit_batch` call (lockstep
    one-CTA-per-fit kernel when a pair has <= 128 rows, sharded over `ngpus`).  Returns
    (rhomatrix, names), directly usable as fitGP(rhomatrix=..., rhomatrix_names=...)."""
    from .batch import fit_batch
    pr = fitGP(data=data, y=y, x=x, pop=pop, time=time, E=E, tau=tau, scaling=scaling, augdata=augdata,
               _prep_only=True)
    up, d = pr["up"], pr["d"]
    n_pop = len(up)
    if n_pop == 1:
        raise ValueError("There must be more than 1 population to get a pairwise matrix.")
    ip = np.concatenate([np.full(d, 0.1), [0.001, 1 - 0.001, 0.5]]) if initpars is None else np.array(initpars, dtype=np.float64)
    if len(ip) != d + 3:
        raise ValueError("initpars must have number of predictors + 3 entries")
    fp = np.full(d + 2, np.nan) if fixedpars is None else np.asarray(fixedpars, dtype=np.float64)
    if len(fp) != d + 2:
        raise ValueError("Length of fixedpars must be number of predictors + 2 (phi, ve, sigma2)")
    fi = ~np.isnan(fp)
    ip[:d + 2][fi] = fp[fi]
    with np.errstate(divide="ignore"):  # :425-426 (rhomin = 0 in the initial transform; two populations: rho stays free)
        initparst = np.concatenate([np.log(ip[:d]), [logit(ip[d], VEMIN, VEMAX), logit(ip[d + 1], SIGMA2MIN, SIGMA2MAX),
                                                     logit(ip[d + 2], 0.0, 1.0)]])
    complete = ~(np.isnan(pr["yds"]) | np.isnan(pr["xds"]).any(axis=1))  # :446 inside each pair's fitGP
    pairs, problems = [], []
    for i in range(n_pop - 1):
        for j in range(i + 1, n_pop):
            rows = np.where(((pr["popd"] == up[i]) | (pr["popd"] == up[j])) & complete)[0]
            pairs.append((i, j))
            problems.append(dict(Y=pr["yds"][rows], X=pr["xds"][rows], pop=pr["popd"][rows], initparst=initparst,
                                 fixedpars=None if fixedpars is None else fp, modeprior=float(modeprior)))
    res = fit_batch(problems, ngpus=ngpus, devices=devices)
    rhomat = np.zeros((n_pop, n_pop))
    for (i, j), r in zip(pairs, res):
        rhomat[i, j] = r["pars"][d + 2]
    return rhomat + rhomat.T + np.eye(n_pop), list(up)
