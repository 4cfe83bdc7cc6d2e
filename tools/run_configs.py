"""Run the BASELINE.json configurations C1-C5 end to end on one B200 and record wall times,
fitted modes and (where the CPU oracle is affordable) parity.  Output: JSON lines."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gpedm_b200 as G
from gpedm_b200 import synth, _lib
from oracle import gpedm_oracle as O

which = sys.argv[1:] or ["C1", "C2", "C3", "C5"]
try:
    from threadpoolctl import threadpool_limits
    threadpool_limits(limits=os.cpu_count())
except Exception:
    pass


def out(**kw):
    print(json.dumps(kw), flush=True)


if "C1" in which:
    import csv
    rows = list(csv.reader(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "thetalog2pop.csv"))))
    tl = {c: (np.array([float(r[j]) for r in rows[1:]]) if c != "Population" else np.array([r[j] for r in rows[1:]], dtype=object))
          for j, c in enumerate(rows[0])}
    G.fitGP(data=tl, y="Abundance", pop="Population", E=3, tau=1, scaling="local", predictmethod="loo").close()
    t = time.perf_counter()
    g = G.fitGP(data=tl, y="Abundance", pop="Population", E=3, tau=1, scaling="local", predictmethod="loo")
    tg = time.perf_counter() - t
    t = time.perf_counter()
    o = O.fitGP(data=tl, y="Abundance", pop="Population", E=3, tau=1, scaling="local", predictmethod="loo")
    to = time.perf_counter() - t
    out(config="C1 thetalog2pop E=3 tau=1 local loo", N=len(g["inputs"]["Y"]), gpu_s=tg, cpu_oracle_s=to, iter=[g["iter"], o["iter"]],
        pars=g["pars"].tolist(), max_mode_err=float(np.max(np.abs(g["pars"] - o["pars"]))),
        loo_R2=[g["outsampfitstats"]["R2"], o["outsampfitstats"]["R2"]])
    g.close()

if "C2" in which:
    y = synth.ricker(4096 + 1024, 1002)
    ytr, yte = y[:4096], y[4096:]
    t = time.perf_counter()
    g = G.fitGP(y=ytr, E=10, tau=1)
    tfit = time.perf_counter() - t
    t = time.perf_counter()
    p = G.predict(g, ynew=yte, xnew=None) if False else G.predict(g, ynew=yte)
    tpred = time.perf_counter() - t
    rec = dict(config="C2 Ricker N=4096 E=10 tau=1 fit + predict 1024 held-out", N=len(g["inputs"]["Y"]), fit_s=tfit,
               predict_s=tpred, iter=g["iter"], nevals=g["nevals"], pars=g["pars"].tolist(),
               R2_in=g["insampfitstats"]["R2"], R2_out=p["outsampfitstats"]["R2"])
    if "--oracle" in sys.argv:
        t = time.perf_counter()
        o = O.fitGP(y=ytr, E=10, tau=1)
        rec["cpu_oracle_fit_s"] = time.perf_counter() - t
        rec["oracle_iter"] = o["iter"]
        rec["max_mode_err"] = float(np.max(np.abs(g["pars"] - o["pars"]) / np.maximum(1, np.abs(o["pars"]))))
        rec["ln_post_rel_err"] = abs(g["insampfitstats"]["ln_post"] - o["insampfitstats"]["ln_post"]) / abs(o["insampfitstats"]["ln_post"])
        po = O.predict(o, ynew=yte)
        rec["predmean_max_abs_err"] = float(np.nanmax(np.abs(p["outsampresults"]["predmean"] - po["outsampresults"]["predmean"])))
        rec["predsd_max_rel_err"] = float(np.nanmax(np.abs(p["outsampresults"]["predsd"] / po["outsampresults"]["predsd"] - 1)))
    rec["predmean"] = p["outsampresults"]["predmean"].tolist()
    rec["predsd"] = p["outsampresults"]["predsd"].tolist()
    rec["ln_post"] = g["insampfitstats"]["ln_post"]
    rec["lnL_LOO"] = g["insampfitstats"]["lnL_LOO"]
    rec["df"] = g["insampfitstats"]["df"]
    out(**rec)
    g.close()

if "C3" in which:
    Y, X, pop = synth.hierarchical_thetalog(npop=16, n=1024, E=5, seed0=1003)
    names = np.array(["pop%02d" % k for k in pop], dtype=object)
    t = time.perf_counter()
    m = G.GPModel(Y, X, names)   # first 6.4 GB of the stream-ordered pool are mapped here (~1.7 s once per process)
    tcreate = time.perf_counter() - t
    from gpedm_b200.batch import default_initparst
    t = time.perf_counter()
    r = m.fit_rprop(default_initparst(5, single_pop=False))
    tfit = time.perf_counter() - t
    t = time.perf_counter()
    mean, var, _ = m.predict_loo(0)
    tloo = time.perf_counter() - t
    rec = dict(config="C3 hierarchical theta-logistic 16 pops x 1024, E=5, shared rho, LOO", N=len(Y), create_s=tcreate, fit_s=tfit, loo_s=tloo,
               iter=r["iter"], nevals=r["nevals"], pars=r["pars"].tolist(), ln_post=-r["nllpost"], lnL_LOO=r["lnL_LOO"],
               df=r["df"], loo_R2=G.getR2(Y, mean), ms_per_eval=1e3 * tfit / r["nevals"])
    # property check: Sigma a = Y on a random sample of rows (recomputed on the host from the fitted modes)
    a = m.vector(_lib.VEC_ALPHA)
    rng = np.random.default_rng(0)
    rows = rng.choice(len(Y), 64, replace=False)
    d = 5
    phi, ve, s2, rho = r["pars"][:d], r["pars"][d], r["pars"][d + 1], r["pars"][d + 2]
    _, Cs = O.getcov(phi, s2, rho, X, X[rows], names, names[rows])
    resid = Cs @ a + ve * a[rows] - Y[rows]
    rec["max_abs_residual_Sigma_a_minus_Y"] = float(np.max(np.abs(resid)))
    # closed-form LOO against the definition on 3 points (one Cholesky solve of size N-1 each, on the host)
    errs = []
    for i in rows[:3]:
        keep = np.delete(np.arange(len(Y)), i)
        _, Cd = O.getcov(phi, s2, rho, X[keep], X[keep], names[keep], names[keep])
        Cd[np.diag_indices(len(keep))] += ve
        _, ci = O.getcov(phi, s2, rho, X[keep], X[i:i + 1], names[keep], names[i:i + 1])
        import scipy.linalg as sl
        cf = sl.cho_factor(Cd, lower=True, overwrite_a=True, check_finite=False)
        sol = sl.cho_solve(cf, np.column_stack([Y[keep], ci[0]]), check_finite=False)
        mu, v = ci[0] @ sol[:, 0], s2 - ci[0] @ sol[:, 1]
        errs.append((abs(mu - mean[i]), abs(np.sqrt(v) / np.sqrt(var[i]) - 1)))
    rec["loo_vs_bruteforce_3pts_mean_abs_sd_rel"] = [[float(a_), float(b_)] for a_, b_ in errs]
    out(**rec)
    m.close()

if "C5" in which:
    t = time.perf_counter()
    hp = synth.hastings_powell(32768 + 3 + 4, 1005, dt=0.05)
    tgen = time.perf_counter() - t
    data = {"Time": np.arange(len(hp), dtype=float), "X": hp[:, 0], "Y": hp[:, 1], "Z": hp[:, 2]}
    lags = G.makelags(data=data, y=["X", "Y", "Z"], E=3, tau=1)
    data.update(lags)
    xcols = list(lags)
    ntr = 32768 + 3
    tr = {k: v[:ntr] for k, v in data.items()}
    te = {k: v[ntr:] for k, v in data.items()}
    t = time.perf_counter()
    g = G.fitGP(data=tr, y="X", x=xcols, time="Time")
    tfit = time.perf_counter() - t
    rec = dict(config="C5 Hastings-Powell 3 species, N=32768, d=9, single large fit", N=len(g["inputs"]["Y"]), gen_s=tgen, fit_s=tfit,
               iter=g["iter"], nevals=g["nevals"], ms_per_eval=1e3 * tfit / g["nevals"], pars=g["pars"].tolist(),
               R2_in=g["insampfitstats"]["R2"], phases_ms=g["model"].phase_ms())
    out(**rec)
    if "--seq" in sys.argv:
        t = time.perf_counter()
        u = G.predict_seq(g, te)
        rec2 = dict(config="C5 predict_seq over 4 appended time steps (4 warm-started refits)", seq_s=time.perf_counter() - t,
                    final_iter=u["iter"], out_R2=u["outsampfitstats"]["R2"], predmean=u["outsampresults"]["predmean"].tolist(),
                    obs=u["outsampresults"]["obs"].tolist())
        out(**rec2)
        u.close()
    g.close()
