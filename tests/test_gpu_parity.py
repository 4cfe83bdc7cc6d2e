"""GPU parity tests: the CUDA path, called through the C ABI (ctypes), against the CPU oracle on
the same inputs, the committed golden fixtures, and size-independent properties at full size.
Tolerances are the north-star ones: nll rel 1e-10, grad 1e-8, fitted modes
1e-6 with identical iteration counts, predmean 1e-9, predfsd rel 1e-9 against the extended-
precision oracle."""
import numpy as np
import pytest

import gpedm_b200 as G
from gpedm_b200 import _lib, batch, synth
from oracle import gpedm_oracle as O
from conftest import arr

pytestmark = pytest.mark.gpu

TOL_NLL, TOL_GRAD, TOL_MODE, TOL_MEAN, TOL_FSD = 1e-10, 1e-8, 1e-6, 1e-9, 1e-9


def grad_err(g, ref):
    return float(np.max(np.abs(g - ref)) / max(1.0, np.max(np.abs(ref))))


def rel(a, b):
    return abs(a - b) / abs(b)


def oracle_eval(parst, Y, X, pop, fixedpars=None, rhofixed=None, rhomatrix=None, names=None, modeprior=1):
    d = X.shape[1]
    fp = np.full(d + 2, np.nan) if fixedpars is None else np.asarray(fixedpars, float)
    D = [O.distance(X[:, i]) for i in range(d)]
    return O.getlikegrad(parst, Y, X, pop, modeprior, fp, rhofixed, rhomatrix, D, False, names)


def make_case(seed, N, d, npop):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((N, d))
    pop = rng.integers(0, npop, N)
    pop[:npop] = np.arange(npop)
    Y = np.sin(X[:, 0]) + 0.3 * np.cos(X[:, -1]) + 0.1 * rng.standard_normal(N)
    parst = np.concatenate([np.log(rng.uniform(0.05, 0.6, d)), [-3.0, 0.5, 0.3]])
    return Y, X, pop, parst


# ------------------------------------------------------------------ getlikegrad
@pytest.mark.parametrize("N,d,npop", [(1, 1, 1), (2, 1, 1), (94, 3, 2), (127, 3, 2), (128, 2, 1), (129, 4, 2), (255, 3, 2),
                                       (256, 5, 1), (257, 2, 3), (300, 4, 1), (512, 8, 2), (641, 6, 4), (700, 5, 3),
                                       (1000, 10, 1), (1301, 32, 5), (2049, 16, 2)])
def test_likegrad_parity(gpu_lib, N, d, npop):
    Y, X, pop, parst = make_case(N + d, N, d, npop)
    ref = oracle_eval(parst, Y, X, pop)
    m = G.GPModel(Y, X, pop)
    go = m.likegrad(parst, 1, None, None, returngradonly=True)
    full = m.likegrad(parst, 1, None, None, returngradonly=False)
    assert go["nllpost"] == full["nllpost"] and np.array_equal(go["grad"], full["grad"])  # deterministic
    assert rel(full["nllpost"], ref["nllpost"]) <= TOL_NLL
    assert grad_err(full["grad"], ref["grad"]) <= TOL_GRAD
    assert rel(full["lnL_LOO"], ref["lnL_LOO"]) <= 1e-9
    assert rel(full["df"], ref["df"]) <= 1e-9
    assert rel(full["SS"], ref["SS"]) <= 1e-9 and rel(full["logdet"], ref["logdet"]) <= 1e-10
    np.testing.assert_allclose(full["pars"], ref["pars"], rtol=1e-15)
    np.testing.assert_allclose(m.vector(_lib.VEC_MEAN), ref["mean"], atol=TOL_MEAN)
    np.testing.assert_allclose(m.vector(_lib.VEC_ALPHA), ref["a"], rtol=1e-7, atol=1e-9)
    m.close()


def test_likegrad_rhomatrix_fixedpars_rhofixed_modeprior(gpu_lib):
    Y, X, pop, parst = make_case(5, 500, 3, 3)
    R = np.array([[1, .3, .5], [.3, 1, .2], [.5, .2, 1.]])
    fp = np.array([np.nan, 0.2, np.nan, 0.01, np.nan])
    for kw in (dict(rhomatrix=R, names=[0, 1, 2]), dict(fixedpars=fp, rhofixed=0.3), dict(modeprior=3),
               dict(fixedpars=np.array([0.1, 0.2, 0.3, 0.02, 1.5]))):
        ref = oracle_eval(parst, Y, X, pop, **kw)
        m = G.GPModel(Y, X, pop, kw.get("rhomatrix"), kw.get("names"))
        out = m.likegrad(parst, kw.get("modeprior", 1), kw.get("fixedpars"), kw.get("rhofixed"), False)
        assert rel(out["nllpost"], ref["nllpost"]) <= TOL_NLL, kw
        assert grad_err(out["grad"], ref["grad"]) <= TOL_GRAD, kw
        np.testing.assert_allclose(out["parst"], ref["parst"], rtol=1e-14)
        m.close()


def test_phi_zero_warm_start_gives_no_nan(gpu_lib):
    """log(phi) = -Inf (underflowed mode passed back as initpars)."""
    Y, X, pop, parst = make_case(9, 200, 3, 1)
    parst[1] = -np.inf
    ref = oracle_eval(parst, Y, X, pop)
    m = G.GPModel(Y, X, pop)
    out = m.likegrad(parst, 1, None, None, True)
    assert np.isfinite(out["nllpost"]) and out["grad"][1] == 0.0 and np.all(np.isfinite(out["grad"]))
    assert rel(out["nllpost"], ref["nllpost"]) <= TOL_NLL and grad_err(out["grad"], ref["grad"]) <= TOL_GRAD
    m.close()


def test_exported_matrices(gpu_lib):
    Y, X, pop, parst = make_case(3, 300, 4, 2)
    ref = oracle_eval(parst, Y, X, pop)
    m = G.GPModel(Y, X, pop)
    m.likegrad(parst, 1, None, None, False)
    np.testing.assert_allclose(m.matrix(_lib.MAT_CD), ref["covm"]["Cd"], rtol=2e-15, atol=1e-300)
    np.testing.assert_allclose(m.matrix(_lib.MAT_SIGMA), ref["covm"]["Sigma"], rtol=2e-15, atol=1e-300)
    np.testing.assert_array_equal(m.matrix(_lib.MAT_POPSAME), ref["covm"]["popsame"])
    S = m.matrix(_lib.MAT_IKVS)
    np.testing.assert_array_equal(S, S.T)
    np.testing.assert_allclose(S, ref["covm"]["iKVs"], rtol=1e-9, atol=1e-9 * np.abs(ref["covm"]["iKVs"]).max())
    Lr = np.linalg.cholesky(ref["covm"]["Sigma"])
    L = m.matrix(_lib.MAT_L)
    assert np.all(np.triu(L, 1) == 0)
    np.testing.assert_allclose(L, Lr, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(m.matrix(_lib.MAT_LINV) @ Lr, np.eye(300), atol=1e-10)
    m.close()


def test_getcov_rectangular(gpu_lib):
    rng = np.random.default_rng(4)
    X, X2 = rng.standard_normal((150, 3)), rng.standard_normal((37, 3))
    pop = np.array(list("abc"), dtype=object)[rng.integers(0, 3, 150)]
    pop2 = np.array(list("abc"), dtype=object)[rng.integers(0, 3, 37)]
    phi = np.array([0.3, 0.0, 1.2])
    ps, Cd = O.getcov(phi, 1.7, 0.4, X, X2, pop, pop2)
    got = G.getcov(phi, 1.7, 0.4, X, X2, pop, pop2)
    np.testing.assert_allclose(got["Cd"], Cd, rtol=2e-15, atol=1e-300)
    np.testing.assert_array_equal(got["popsame"], ps)
    R = np.array([[1, .1, .2], [.3, 1, .4], [.5, .6, 1.]])  # deliberately asymmetric: [pop2, pop] addressing
    names = unique = O.unique_stable(pop)
    _, Cd = O.getcov(phi, 1.7, 0.4, X, X2, pop, pop2, R, names)
    got = G.getcov(phi, 1.7, 0.4, X, X2, pop, pop2, R, names)
    np.testing.assert_allclose(got["Cd"], Cd, rtol=2e-15, atol=1e-300)


def test_not_positive_definite_is_reported(gpu_lib):
    Y, X, pop, parst = make_case(8, 200, 2, 1)
    X[57, 0] = np.nan  # poisons row/column 58 of Sigma
    m = G.GPModel(Y, X, pop)
    with pytest.raises(G.GPEDMError) as e:
        m.likegrad(parst, 1, None, None, True)
    assert e.value.status >= 1 and "not positive definite" in str(e.value)
    m.close()


# ------------------------------------------------------------------ README cases through fitGP / predict
def check_fit(g, ref, d):
    assert g["iter"] == ref["iter"]
    pars, rp = g["pars"], arr(ref["pars"])
    assert np.max(np.abs(pars - rp) / np.maximum(1, np.abs(rp))) <= TOL_MODE
    np.testing.assert_allclose(g["insampresults"]["predmean"], arr(ref["insampresults"]["predmean"]), atol=1e-8)
    np.testing.assert_allclose(g["insampresults"]["predsd"], arr(ref["insampresults"]["predsd"]), rtol=1e-7)
    for k in ("R2", "rmse", "ln_post", "ln_prior", "lnL_LOO", "df", "SS", "logdet"):
        assert g["insampfitstats"][k] == pytest.approx(ref["insampfitstats"][k], rel=1e-7), k


def check_pred(p, ref, mean_tol=TOL_MEAN, sd_rel=1e-7):
    a, b = p["outsampresults"], ref["outsampresults"]
    np.testing.assert_allclose(a["predmean"], arr(b["predmean"]), atol=mean_tol, equal_nan=True)
    np.testing.assert_allclose(a["predsd"], arr(b["predsd"]), rtol=sd_rel, equal_nan=True)
    np.testing.assert_allclose(a["predfsd"], arr(b["predfsd"]), rtol=1e-6, equal_nan=True)
    if "GPgrad" in ref:
        np.testing.assert_allclose(p["GPgrad"], np.array([arr(r) for r in ref["GPgrad"]]), atol=1e-8, equal_nan=True)
    if "outsampfitstats" in ref:
        assert p["outsampfitstats"]["R2"] == pytest.approx(ref["outsampfitstats"]["R2"], rel=1e-8)


@pytest.fixture(scope="module")
def k1(gpu_lib, thetalog2pop):
    g = G.fitGP(data=thetalog2pop, y="Abundance", pop="Population", E=3, tau=1, scaling="local", predictmethod="loo")
    yield g
    g.close()


def test_K1_fit_and_loo(k1, oracle_cases):
    ref = oracle_cases["K1"]
    check_fit(k1, ref, 3)
    check_pred(k1, ref)
    # README.md:79-101 printed digits
    assert round(k1["pars"][0], 4) == 0.5317 and float("%.7g" % k1["pars"][3]) == 0.01222333
    assert float("%.7g" % k1["outsampfitstats"]["R2"]) == 0.991238
    assert float("%.7g" % k1["outsampfitstatspop"]["R2pop"]["PopB"]) == 0.9754702


@pytest.mark.parametrize("key,kw", [("lto", dict(predictmethod="lto")),
                                    ("lto_r1", dict(predictmethod="lto", exclradius=1)),
                                    ("loo_r2", dict(predictmethod="loo", exclradius=2)),
                                    ("loo_grad", dict(predictmethod="loo", returnGPgrad=True)),
                                    ("lto_grad", dict(predictmethod="lto", returnGPgrad=True)),
                                    ("sequential", dict(predictmethod="sequential"))])
def test_K1_predict_methods(k1, oracle_cases, key, kw):
    check_pred(G.predict(k1, **kw), oracle_cases["K1"][key])


def test_K1_covm_is_lazy_and_matches(k1):
    assert not dict.__contains__(k1, "covm")
    covm = k1["covm"]
    n = len(k1["inputs"]["Y"])
    assert covm["iKVs"].shape == (n, n)
    np.testing.assert_allclose(covm["Sigma"] - covm["Cd"], k1["pars"][3] * np.eye(n), atol=1e-15)
    np.testing.assert_allclose(covm["Sigma"] @ covm["iKVs"], np.eye(n), atol=1e-8)


def test_K5(gpu_lib, thetalog2pop, oracle_cases):
    g = G.fitGP(y=thetalog2pop["Abundance"], pop=thetalog2pop["Population"], E=2, tau=1, scaling="local")
    check_fit(g, oracle_cases["K5"], 2)
    assert round(g["pars"][0], 5) == 0.53026 and float("%.7g" % g["pars"][4]) == 0.3250384  # README.md:557-571
    g.close()


def _popA(tl):
    pA = {k: v[tl["Population"] == "PopA"] for k, v in tl.items()}
    lags = G.makelags(data=pA, y="Abundance", E=2, tau=1)
    pA.update(lags)
    return {k: v[:40] for k, v in pA.items()}, {k: v[40:50] for k, v in pA.items()}


def test_K2_newdata_sequential_and_K3_predict_seq(gpu_lib, thetalog2pop, oracle_cases):
    tr, te = _popA(thetalog2pop)
    g = G.fitGP(data=tr, y="Abundance", x=["Abundance_1", "Abundance_2"], time="Time", newdata=te)
    check_fit(g, oracle_cases["K2"], 2)
    check_pred(g, oracle_cases["K2"])
    check_pred(G.predict(g, predictmethod="sequential"), oracle_cases["K2"]["sequential"])
    assert round(g["pars"][0], 5) == 0.62609 and float("%.7g" % g["pars"][2]) == 0.003394526  # README.md:291-300
    u = G.predict_seq(g, te)
    ref = oracle_cases["K3"]
    check_fit(u, ref, 2)
    check_pred(u, ref)
    assert round(u["pars"][0], 5) == 0.59103 and float("%.7g" % u["outsampfitstats"]["R2"]) == 0.9973224  # :302-312
    g.close()
    u.close()


def test_K4_forecast_with_gradient(gpu_lib, thetalog2pop, oracle_cases):
    tl = thetalog2pop
    d1 = G.makelags(data=tl, y="Abundance", pop="Population", time="Time", E=3, tau=1, append=True)
    fore = G.makelags(data=tl, y="Abundance", pop="Population", time="Time", E=3, tau=1, forecast=True)
    np.testing.assert_allclose(np.column_stack([fore["Abundance_%d" % i] for i in (1, 2, 3)]),
                               np.array(oracle_cases["K4"]["fore1"]), rtol=1e-15)
    g = G.fitGP(data=d1, y="Abundance", x=["Abundance_1", "Abundance_2", "Abundance_3"], pop="Population",
                time="Time", scaling="local")
    check_fit(g, oracle_cases["K4"], 3)
    p = G.predict(g, newdata=fore, returnGPgrad=True)
    check_pred(p, oracle_cases["K4"]["forecast"])
    r = p["outsampresults"]  # README.md:444-452
    assert ["%.7g" % v for v in r["predmean"]] == ["0.04932054", "1.416254"]
    assert ["%.7g" % v for v in r["predfsd"]] == ["0.01647408", "0.01679414"]
    g.close()


# ------------------------------------------------------------------ variance accuracy
def test_predfsd_against_extended_precision(k1):
    """predfsd rel error <= 1e-9 against the 80-bit oracle; the FP64 restatement of the reference's own
    formula (sigma2 - c' iKVs c) is itself ~1e-8..1e-6 away from it."""
    d = 3
    p = k1["pars"]
    X, Y, Pop = k1["inputs"]["X"], k1["inputs"]["Y"], k1["inputs"]["Pop"]
    rng = np.random.default_rng(2)
    Xn = X[:40] + 0.05 * rng.standard_normal((40, d))
    Pn = Pop[:40]
    mean, var, _ = k1["model"].predict_new(Xn, Pn)
    mt, vt = O.predict_newdata_ld(p[:d], p[d + 1], p[d + 2], p[d], X, Y, Pop, Xn, Pn)
    mt, vt = np.asarray(mt, float), np.asarray(vt, float)
    assert np.max(np.abs(mean - mt)) <= TOL_MEAN
    err = np.max(np.abs(np.sqrt(var) / np.sqrt(vt) - 1))
    assert err <= TOL_FSD, err
    _, v64, _ = O.predict_newdata_core(p[:d], p[d + 1], p[d + 2], p[d], X, Y, Pop, k1["covm"]["iKVs"], Xn, Pn)
    print("predfsd rel err vs 80-bit: gpu %.2e, fp64 reference formula %.2e" %
          (err, np.max(np.abs(np.sqrt(v64) / np.sqrt(vt) - 1))))
    # closed-form LOO variance against the extended-precision inverse diagonal
    _, _, dinv = O.loglik_ld(p[:d], p[d + 1], p[d + 2], p[d], X, Y, Pop)
    loov = k1["model"].vector(_lib.VEC_LOO_VAR)
    truth = np.asarray(1 / dinv - p[d], float)
    assert np.max(np.abs(np.sqrt(loov) / np.sqrt(truth) - 1)) <= TOL_FSD


# ------------------------------------------------------------------ other front-end paths
def test_augdata_loo_with_duplicates(gpu_lib):
    rng = np.random.default_rng(12)
    y = synth.ricker(60, 3)
    lag = np.concatenate([[np.nan], y[:-1]])
    data = {"t": np.arange(60.0), "y": y, "y_1": lag}
    aug = {"t": np.array([10., 20., 30.]), "y": y[[10, 20, 30]], "y_1": lag[[10, 20, 30]] + 0.01 * rng.standard_normal(3)}
    kw = dict(y="y", x="y_1", time="t", augdata=aug, predictmethod="loo", exclradius=1)
    g = G.fitGP(data=data, **kw)
    o = O.fitGP(data=data, **kw)
    assert g["iter"] == o["iter"]
    np.testing.assert_allclose(g["pars"], o["pars"], rtol=1e-6, atol=1e-300)
    np.testing.assert_allclose(g["outsampresults"]["predmean"], o["outsampresults"]["predmean"], atol=1e-8, equal_nan=True)
    np.testing.assert_allclose(g["outsampresults"]["predsd"], o["outsampresults"]["predsd"], rtol=1e-6, equal_nan=True)
    g.close()


def test_hierarchical_rhomatrix_and_fixed(gpu_lib):
    Y, X, pop = synth.hierarchical_thetalog(npop=4, n=60, E=2, seed0=40)
    names = ["p%d" % k for k in pop]
    data = {"y": Y, "x1": X[:, 0], "x2": X[:, 1], "pop": np.array(names, dtype=object)}
    R = 0.3 + 0.7 * np.eye(4)
    for kw in (dict(rhomatrix=R, rhomatrix_names=["p0", "p1", "p2", "p3"]), dict(rhofixed=0.25),
               dict(fixedpars=[np.nan, 0.0, 0.01, np.nan]), dict(modeprior=2, scaling="local")):
        g = G.fitGP(data=data, y="y", x=["x1", "x2"], pop="pop", predictmethod="loo", **kw)
        o = O.fitGP(data=data, y="y", x=["x1", "x2"], pop="pop", predictmethod="loo", **kw)
        assert g["iter"] == o["iter"], kw
        assert np.max(np.abs(g["pars"] - o["pars"]) / np.maximum(1, np.abs(o["pars"]))) <= TOL_MODE, kw
        np.testing.assert_allclose(g["outsampresults"]["predmean"], o["outsampresults"]["predmean"], atol=1e-8)
        g.close()


def test_batch_equals_individual_fits(gpu_lib):
    y = synth.ricker(260, 21)
    probs, meta = batch.grid_problems(y, [1, 2, 3], [1, 2])
    res = G.fit_batch(probs, ngpus=1, want_loo=True)
    for pr, r in zip(probs, res):
        m = G.GPModel(pr["Y"], pr["X"])
        one = m.fit_rprop(batch.default_initparst(pr["X"].shape[1]))
        assert one["iter"] == r["iter"] and one["nllpost"] == r["nllpost"]
        np.testing.assert_array_equal(one["pars"], r["pars"])
        np.testing.assert_array_equal(m.vector(_lib.VEC_LOO_MEAN), r["loo_mean"])
        m.close()
    grid = G.fit_grid(y, [1, 2], [1], ngpus=1)
    assert [g["E"] for g in grid] == [1, 2] and all(0.5 < g["R2_loo"] <= 1 for g in grid)
    o = O.fitGP(y=y, E=2, tau=1, predictmethod="loo")
    assert grid[1]["iter"] == o["iter"]
    assert grid[1]["R2_loo"] == pytest.approx(o["outsampfitstats"]["R2"], rel=1e-7)


# ------------------------------------------------------------------ full-size properties (BASELINE sizes)
def test_full_size_properties_N8182(gpu_lib):
    """At the headline size (N=8182, d=10) the oracle's O(N^2 d) temporaries are too slow for a unit
    test, so check size-independent properties: (1) nll against a plain LAPACK Cholesky of the same
    Sigma; (2) Sigma a = Y; (3) gradient vs central finite differences of the GPU's own nll for ve and
    sigma2 (phi carries the reference's one-half quirk and is checked at small N); (4) determinism."""
    Y, X = synth.embed(synth.ricker(8192, 1004), 10, 1)
    N = len(Y)
    p0 = batch.default_initparst(10, True)
    p0[:10] = np.log([0.3, 0.1, 0.05, 0.02, 0.01, 0.01, 0.01, 0.01, 0.01, 0.01])
    p0[10] = -4.0
    m = G.GPModel(Y, X)
    r = m.likegrad(p0, 1, None, None, False)
    r2 = m.likegrad(p0, 1, None, None, False)
    assert r["nllpost"] == r2["nllpost"] and np.array_equal(r["grad"], r2["grad"])
    a = m.vector(_lib.VEC_ALPHA)
    diag_s = m.vector(_lib.VEC_DIAG_IKVS)
    phi, ve, s2 = r["pars"][:10], r["pars"][10], r["pars"][11]
    Xs = X * np.sqrt(phi)
    sq = (Xs ** 2).sum(1)
    Sigma = s2 * np.exp(-np.maximum(sq[:, None] + sq[None, :] - 2 * Xs @ Xs.T, 0))
    Sigma[np.diag_indices(N)] = s2 + ve
    np.testing.assert_allclose(Sigma @ a, Y, atol=2e-7)
    L = np.linalg.cholesky(Sigma)
    z = np.linalg.solve(L, Y) if N < 3000 else __import__("scipy.linalg", fromlist=["x"]).solve_triangular(L, Y, lower=True)
    like = -0.5 * z @ z - np.log(np.diag(L)).sum()
    assert rel(r["like"], like) <= 1e-9  # Sigma above is built with a different (GEMM) distance formula
    h = 1e-5
    for k in (10, 11):
        pp, pm = p0.copy(), p0.copy()
        pp[k] += h
        pm[k] -= h
        fd = (m.likegrad(pp, 1, None, None, True)["nllpost"] - m.likegrad(pm, 1, None, None, True)["nllpost"]) / (2 * h)
        assert fd == pytest.approx(r["grad"][k], rel=2e-5)
    assert abs(r["df"] - (N - ve * np.sum(diag_s))) < 1e-6
    with pytest.raises(G.GPEDMError):  # gradient-only evaluations invalidate the resident factorisation
        m.vector(_lib.VEC_DIAG_IKVS)
    m.close()


def test_phi_gradient_half_quirk(gpu_lib):
    """The reference's phi gradient is 0.5 * d(like)/d(phi) + d(lp)/d(phi) (R/GPfunctions.R:663): finite
    differences of like and lp must reproduce the GPU gradient only with that one half."""
    Y, X, pop, parst = make_case(77, 400, 2, 2)
    m = G.GPModel(Y, X, pop)
    r = m.likegrad(parst, 1, None, None, True)
    h = 1e-6
    for k in range(2):
        pp, pm = parst.copy(), parst.copy()
        pp[k] += h
        pm[k] -= h
        a, b = m.likegrad(pp, 1, None, None, True), m.likegrad(pm, 1, None, None, True)
        dlike = (a["like"] - b["like"]) / (2 * h)
        dlp = (a["lp"] - b["lp"]) / (2 * h)
        assert -(0.5 * dlike + dlp) == pytest.approx(r["grad"][k], rel=1e-5)
    m.close()


def test_batch_two_gpus_nccl_gather(gpu_lib):
    """In-library multi-GPU batch: worker thread per GPU + ONE ncclAllGather of the result records."""
    if gpu_lib.gpedm_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    y = synth.ricker(400, 5)
    probs, _ = batch.grid_problems(y, [1, 2, 3, 4], [1, 2])
    one = G.fit_batch(probs, ngpus=1)
    two = G.fit_batch(probs, ngpus=2)
    for a, b in zip(one, two):
        assert a["iter"] == b["iter"] and a["nllpost"] == b["nllpost"] and np.array_equal(a["pars"], b["pars"])
    # the grid entry point (series uploaded once per GPU, device-side preparation) and a batch of
    # one-tile fits (lockstep kernel, a strided share per GPU) through the same record exchange
    g1 = G.fit_grid(y, [1, 2, 3], [1, 2], ngpus=1)
    g2 = G.fit_grid(y, [1, 2, 3], [1, 2], ngpus=2)
    for a, b in zip(g1, g2):
        assert (a["E"], a["tau"], a["iter"], a["N"]) == (b["E"], b["tau"], b["iter"], b["N"])
        assert a["ln_post"] == b["ln_post"] and a["R2_loo"] == b["R2_loo"] and np.array_equal(a["pars"], b["pars"])
    small, _ = batch.grid_problems(synth.ricker(100, 6), [1, 2, 3], [1, 2])
    s1 = G.fit_batch(small, ngpus=1, want_loo=True)
    s2 = G.fit_batch(small, ngpus=2, want_loo=True)
    for a, b in zip(s1, s2):
        assert a["iter"] == b["iter"] and a["nllpost"] == b["nllpost"] and np.array_equal(a["loo_mean"], b["loo_mean"])


def test_predict_iter_and_getconditionals(gpu_lib, thetalog2pop):
    """'next' row of SURVEY section 8f: predict_iter (R/predict_iter.R:57-163) and getconditionals
    (R/PlotsConditionals.R:166-266) on the resident factorisation, against the oracle and the README's
    iterated-forecast rows (README.md:444-456)."""
    tl = thetalog2pop
    d1 = G.makelags(data=tl, y="Abundance", pop="Population", time="Time", E=3, tau=1, append=True)
    fore = G.makelags(data=tl, y="Abundance", pop="Population", time="Time", E=3, tau=1, forecast=True)
    cols = ["Abundance_1", "Abundance_2", "Abundance_3"]
    kw = dict(y="Abundance", x=cols, pop="Population", time="Time", scaling="local")
    g = G.fitGP(data=d1, **kw)
    o = O.fitGP(data=d1, **kw)
    nd = {"Time": np.array([51., 51., 52., 52., 53., 53., 54., 54.]), "Population": np.array(["PopA", "PopB"] * 4, dtype=object)}
    for c in cols:
        nd[c] = np.concatenate([fore[c], np.full(6, np.nan)])
    rg = G.predict_iter(g, nd)["outsampresults"]
    ro = O.predict_iter(o, nd)["outsampresults"]
    np.testing.assert_allclose(rg["predmean"], ro["predmean"], atol=1e-8)
    np.testing.assert_allclose(rg["predsd"], ro["predsd"], rtol=1e-6)
    assert ["%.7g" % v for v in rg["predmean"][:6]] == ["0.04932054", "1.416254", "0.15775", "0.265825", "0.5336754", "0.8318641"]
    cg = G.getconditionals(g)
    co = O.getconditionals(o)
    assert len(cg) == 2
    for a, b in zip(cg, co):
        np.testing.assert_allclose(a["xval"], b["xval"], rtol=1e-13)
        np.testing.assert_allclose(a["predmean"], b["predmean"], atol=1e-8)
        np.testing.assert_allclose(a["predsd"], b["predsd"], rtol=1e-5, atol=1e-9)
    cg2 = G.getconditionals(g, xrange="global", nvals=7)
    co2 = O.getconditionals(o, xrange="global", nvals=7)
    np.testing.assert_allclose(cg2[1]["predmean"], co2[1]["predmean"], atol=1e-8)
    g.close()


def test_getcovinv(gpu_lib):
    """getcovinv (R/GPfunctions.R:818-830) on caller-supplied matrices, incl. a slightly asymmetric one
    (the reference symmetrises before factoring) and a non-PD one."""
    rng = np.random.default_rng(6)
    for n in (5, 130, 300):
        B = rng.standard_normal((n, n))
        S = B @ B.T / n + 0.5 * np.eye(n)
        S[0, n - 1] += 1e-13
        out = G.getcovinv(S)
        Ssym = (S + S.T) / 2
        Lr, ir = O.getcovinv(Ssym)
        np.testing.assert_allclose(out["L"], Lr, rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(out["iKVs"], ir, rtol=1e-9, atol=1e-10)
    with pytest.raises(G.GPEDMError) as e:
        G.getcovinv(np.array([[1.0, 2.0], [2.0, 1.0]]))
    assert e.value.status == 2


def test_predict_new_chunking_and_set_data(gpu_lib):
    """More new points than one prediction chunk (4096), and re-uploading the training set in place."""
    Y, X, pop, parst = make_case(21, 260, 3, 2)
    m = G.GPModel(Y, X, pop)
    full = m.likegrad(parst, 1, None, None, False)
    rng = np.random.default_rng(3)
    Xn = rng.standard_normal((4500, 3))
    pn = rng.integers(0, 2, 4500)
    mean, var, grad = m.predict_new(Xn, pn, returnGPgrad=True)
    ref = oracle_eval(parst, Y, X, pop)
    p = ref["pars"]
    mo, vo, go = O.predict_newdata_core(p[:3], p[4], p[5], p[3], X, Y, pop, ref["covm"]["iKVs"], Xn, pn, returnGPgrad=True)
    np.testing.assert_allclose(mean, mo, atol=TOL_MEAN)
    np.testing.assert_allclose(grad, go, atol=1e-8)
    np.testing.assert_allclose(np.sqrt(var), np.sqrt(vo), rtol=1e-6)
    # same handle, new data of the same shape
    Y2, X2, pop2, _ = make_case(22, 260, 3, 2)
    codes2 = np.array([m.levels.index(q) for q in pop2], dtype=np.int32)
    m.set_data(Y2, X2, codes2)
    with pytest.raises(G.GPEDMError):
        m.vector(_lib.VEC_ALPHA)  # the resident factorisation belonged to the old data
    out = m.likegrad(parst, 1, None, None, True)
    ref2 = oracle_eval(parst, Y2, X2, pop2)
    assert rel(out["nllpost"], ref2["nllpost"]) <= TOL_NLL and grad_err(out["grad"], ref2["grad"]) <= TOL_GRAD
    m.close()


def test_small_fit_batch_matches_tiled_path_and_oracle(gpu_lib, thetalog2pop):
    """Batches whose fits all have N <= 128 take the one-CTA-per-fit lockstep path;
    it must agree with the tiled per-handle path and with the oracle (K1 is such a fit)."""
    tl = thetalog2pop
    probs = []
    # K1 itself (N=94, 2 pops, local scaling) built by the fitGP front end of the oracle
    o = O.fitGP(data=tl, y="Abundance", pop="Population", E=3, tau=1, scaling="local")
    probs.append(dict(Y=o["inputs"]["Y"], X=o["inputs"]["X"], pop=o["inputs"]["Pop"],
                      initparst=batch.default_initparst(3, single_pop=False)))
    # an E x tau grid on a short single-population series
    y = synth.ricker(120, 8)
    g, _ = batch.grid_problems(y, [1, 2, 3, 4], [1, 2])
    probs += g
    # pairwise fits with a fixed rho and a fixed ve (getrhomatrix-style clients)
    Y, X, pop = synth.hierarchical_thetalog(npop=3, n=44, E=2, seed0=70)
    for a, b in [(0, 1), (0, 2), (1, 2)]:
        m = (pop == a) | (pop == b)
        probs.append(dict(Y=Y[m], X=X[m], pop=pop[m], initparst=batch.default_initparst(2, single_pop=False)))
    probs.append(dict(Y=Y, X=X, pop=pop, initparst=batch.default_initparst(2, False), rhofixed=0.3,
                      fixedpars=[np.nan, np.nan, 0.01, np.nan]))
    assert max(len(p["Y"]) for p in probs) <= 128
    res = G.fit_batch(probs, ngpus=1, want_loo=True)
    for pr, r in zip(probs, res):
        m = G.GPModel(pr["Y"], pr["X"], pr.get("pop"))
        one = m.fit_rprop(pr.get("initparst", batch.default_initparst(np.atleast_2d(pr["X"].T).T.shape[1])), 1.0,
                          pr.get("fixedpars"), pr.get("rhofixed"))
        assert one["iter"] == r["iter"] and r["status"] == 0
        assert np.max(np.abs(one["pars"] - r["pars"]) / np.maximum(1, np.abs(one["pars"]))) <= 1e-9
        assert rel(r["nllpost"], one["nllpost"]) <= 1e-11
        assert rel(r["lnL_LOO"], one["lnL_LOO"]) <= 1e-9 and rel(r["df"], one["df"]) <= 1e-9
        np.testing.assert_allclose(r["loo_mean"], m.vector(_lib.VEC_LOO_MEAN), atol=1e-9)
        np.testing.assert_allclose(r["loo_var"], m.vector(_lib.VEC_LOO_VAR), rtol=1e-6)
        np.testing.assert_allclose(r["insamp_mean"], m.vector(_lib.VEC_MEAN), atol=1e-9)
        m.close()
    # K1 against the oracle (README.md:79-101)
    assert res[0]["iter"] == o["iter"] == 38
    assert np.max(np.abs(res[0]["pars"] - o["pars"]) / np.maximum(1, np.abs(o["pars"]))) <= TOL_MODE
    assert rel(res[0]["nllpost"], o["nllpost"]) <= TOL_NLL


def test_mixed_batch_small_and_tiled(gpu_lib):
    """A batch mixing one-tile fits (lockstep small-fit kernel) and larger ones (tiled pipeline, worker
    threads) must return, per fit, what an individual fit returns."""
    probs = []
    y = synth.ricker(110, 21)
    g, _ = batch.grid_problems(y, [1, 2, 3], [1])
    probs += g
    for n, E, seed in [(300, 3, 22), (520, 4, 23)]:
        Yb, Xb = synth.embed(synth.ricker(n, seed), E, 1)
        probs.append(dict(Y=Yb, X=Xb, initparst=batch.default_initparst(E, True)))
    sizes = [len(p["Y"]) for p in probs]
    assert min(sizes) <= 128 < max(sizes)
    res = G.fit_batch(probs, ngpus=1, want_loo=True)
    for pr, r in zip(probs, res):
        m = G.GPModel(pr["Y"], pr["X"], pr.get("pop"))
        one = m.fit_rprop(pr.get("initparst", batch.default_initparst(np.atleast_2d(pr["X"].T).T.shape[1])), 1.0,
                          None, None)
        assert r["status"] == 0 and one["iter"] == r["iter"]
        assert np.max(np.abs(one["pars"] - r["pars"]) / np.maximum(1, np.abs(one["pars"]))) <= 1e-9
        assert rel(r["nllpost"], one["nllpost"]) <= 1e-11
        np.testing.assert_allclose(r["loo_mean"], m.vector(_lib.VEC_LOO_MEAN), atol=1e-9)
        m.close()


def test_leaveout_blocks_on_device_vs_bruteforce(gpu_lib):
    """lto / loo with wide exclusion windows (blocks of Sigma^-1 up to ~60 rows, solved by
    leaveout_blocks_kernel) against the oracle's brute-force refactorisation loops
    (reference R/GPfunctions.R:1056-1104, :1140-1194)."""
    Y, X, pop = synth.hierarchical_thetalog(npop=4, n=40, E=2, seed0=90)
    N = len(Y)
    time = np.concatenate([np.arange(np.sum(pop == k)) for k in range(4)]).astype(float)
    parst = np.array([np.log(0.4), np.log(0.05), -3.0, 0.3, 0.2])
    m = G.GPModel(Y, X, pop)
    full = m.likegrad(parst, 1.0, None, None, returngradonly=False)
    D = [O.distance(X[:, i]) for i in range(2)]
    ref = O.getlikegrad(parst, Y, X, pop, 1.0, np.full(4, np.nan), None, None, D, False)
    assert rel(full["nllpost"], ref["nllpost"]) <= TOL_NLL
    primary = np.ones(N, dtype=bool)
    sigma2, ve = ref["pars"][3], ref["pars"][2]
    Cd, Sigma = ref["covm"]["Cd"], ref["covm"]["Sigma"]
    for method, r in [("lto", 0), ("lto", 7), ("loo", 3), ("loo", 12)]:
        if method == "lto":
            gm, gv, _ = m.predict_lto(time, r, None)
            om, ov, _ = O.predict_lto_core(sigma2, Cd, Sigma, Y, time, primary, r)
        else:
            gm, gv, _ = m.predict_loo(r, time, None)
            om, ov, _ = O.predict_loo_core(sigma2, Cd, Sigma, Y, pop, time, primary, r)
        ok = ~np.isnan(om)
        assert np.array_equal(np.isnan(gm), np.isnan(om))
        assert np.max(np.abs(gm[ok] - om[ok])) <= 1e-9
        assert np.max(np.abs(np.sqrt(gv[ok] + ve) - np.sqrt(ov[ok] + ve))) <= 1e-7
    m.close()


# ------------------------------------------------------------------ on-device data preparation
def _host_prep(y, popv, E, tau, scaling):
    """fitGP's lagging / scaling / complete-row selection on the host (gpedm_b200.fitgp, itself pinned to
    the oracle and the README fits)."""
    from gpedm_b200.fitgp import _lagmatrix, _mean, _sd, unique_stable
    x = _lagmatrix(y, popv, E, tau)
    cols = np.column_stack([y, x])
    up = unique_stable(popv)
    G_ = len(up) if scaling == "local" else 1
    center, scale = np.zeros((E + 1, G_)), np.ones((E + 1, G_))
    sc = cols.copy()
    if scaling == "global":
        for j in range(E + 1):
            center[j, 0], scale[j, 0] = _mean(cols[:, j]), _sd(cols[:, j])
        sc = (cols - center[:, 0][None, :]) / scale[:, 0][None, :]
    elif scaling == "local":
        for g, u in enumerate(up):
            m = popv == u
            for j in range(E + 1):
                center[j, g], scale[j, g] = _mean(cols[m, j]), _sd(cols[m, j])
            sc[m] = (cols[m] - center[:, g][None, :]) / scale[:, g][None, :]
    ok = ~np.isnan(sc).any(axis=1)
    return sc[ok, 0], sc[ok, 1:], np.where(ok)[0], center, scale


@pytest.mark.gpu
@pytest.mark.parametrize("scaling,npop,E,tau", [("global", 1, 3, 1), ("none", 1, 2, 2), ("local", 3, 2, 1),
                                                ("global", 2, 4, 3), ("local", 2, 1, 1)])
def test_create_lagged_matches_host_preparation(gpu_lib, scaling, npop, E, tau):
    rng = np.random.default_rng(5 + E)
    ys, pops = [], []
    for k in range(npop):
        ys.append(synth.thetalogistic(70 + 9 * k, 30 + k) * (1 + k))
        pops += ["p%d" % k] * (70 + 9 * k)
    y = np.concatenate(ys)
    y[rng.choice(len(y), 4, replace=False)] = np.nan  # missing observations
    popv = np.asarray(pops, dtype=object)
    if npop > 1:  # interleave the populations row-wise: lags must follow each population's own order
        seq = rng.permutation(popv)
        nxt = {u: list(np.where(popv == u)[0]) for u in set(pops)}
        perm = np.array([nxt[u].pop(0) for u in seq])
        y, popv = y[perm], popv[perm]
    Yh, Xh, rows, center, scale = _host_prep(y, popv, E, tau, scaling)
    m = G.GPModel.from_series(y, E, tau, pop=popv, scaling=scaling)
    assert m.N == len(Yh) and np.array_equal(m.rows, rows)
    if scaling != "none":
        np.testing.assert_allclose(m.center, center, rtol=1e-14, atol=1e-15)
        np.testing.assert_allclose(m.scale, scale, rtol=1e-13)
    np.testing.assert_allclose(m.Y, Yh, rtol=0, atol=1e-12)
    np.testing.assert_allclose(m.X, Xh, rtol=0, atol=1e-12)
    # a fit on the device-prepared set equals a fit on the host-prepared set
    p0 = batch.default_initparst(E, npop == 1)
    r = m.fit_rprop(p0)
    codes = G.core.pop_codes(popv)[0][rows]  # scaling groups are indexed in full-series order of appearance
    mh = G.GPModel(Yh, Xh, popv[rows])
    rh = mh.fit_rprop(p0)
    assert r["iter"] == rh["iter"]
    assert np.max(np.abs(r["pars"] - rh["pars"]) / np.maximum(1, np.abs(rh["pars"]))) <= 1e-8
    assert rel(r["nllpost"], rh["nllpost"]) <= 1e-10
    # device fit statistics vs getR2 on back-transformed host vectors
    from gpedm_b200.fitgp import getR2
    g = np.zeros(len(rows), dtype=int) if scaling != "local" else codes
    c0, s0 = (center[0, g], scale[0, g]) if scaling != "none" else (0.0, 1.0)
    obs = Yh * s0 + c0
    st = m.fit_stats()
    assert st["R2_insample"] == pytest.approx(getR2(obs, mh.vector(_lib.VEC_MEAN) * s0 + c0), rel=1e-9)
    assert st["R2_loo"] == pytest.approx(getR2(obs, mh.vector(_lib.VEC_LOO_MEAN) * s0 + c0), rel=1e-9)
    assert st["n"] == len(rows)
    m.close()
    mh.close()


@pytest.mark.gpu
def test_K1_from_raw_series_on_device(gpu_lib, thetalog2pop, oracle_cases):
    """K1 (README.md:79-101) with the whole data preparation on the device: raw Abundance + Population
    -> lags E=3 tau=1, local scaling, complete rows -> fit -> modes and LOO R2 of the README."""
    tl = thetalog2pop
    m = G.GPModel.from_series(np.asarray(tl["Abundance"], dtype=float), 3, 1, pop=np.asarray(tl["Population"], dtype=object),
                              scaling="local")
    assert m.N == 94
    r = m.fit_rprop(batch.default_initparst(3, single_pop=False))
    o = oracle_cases["K1"]
    assert r["iter"] == o["iter"] == 38
    assert np.max(np.abs(r["pars"] - np.asarray(o["pars"])) / np.maximum(1, np.abs(o["pars"]))) <= TOL_MODE
    st = m.fit_stats()
    assert st["R2_loo"] == pytest.approx(0.991238, abs=5e-7)        # README.md:99
    assert st["R2_insample"] == pytest.approx(0.9933975, abs=5e-8)  # README.md:92
    m.close()


@pytest.mark.gpu
def test_fit_grid_device_prep_equals_host_prep(gpu_lib):
    y = synth.ricker(300, 33)
    dev = G.fit_grid(y, [1, 2, 3], [1, 2], ngpus=1)
    host = batch.fit_grid_hostprep(y, [1, 2, 3], [1, 2], ngpus=1)
    assert [(a["E"], a["tau"]) for a in dev] == [(b["E"], b["tau"]) for b in host]
    for a, b in zip(dev, host):
        assert a["status"] == 0 and a["iter"] == b["iter"] and a["N"] == 300 - a["E"] * a["tau"]
        assert np.max(np.abs(a["pars"] - b["pars"]) / np.maximum(1, np.abs(b["pars"]))) <= 1e-8
        assert a["R2_loo"] == pytest.approx(b["R2_loo"], rel=1e-9)
        assert a["R2_insample"] == pytest.approx(b["R2_insample"], rel=1e-9)


def test_fit_grid_lockstep_statistics_match_single_handles_and_oracle(gpu_lib):
    """gpedm_fit_grid on a short series runs its cells through the lockstep batch and reduces the fit statistics on the
    host from the returned in-sample / leave-one-out means (fit_grid_lockstep, host_lagged.cuh); a single handle
    reduces them on the device (fit_stats_kernel).  Two populations with local scaling exercise the per-population
    unscaling; the oracle's fitGP (reference R/GPfunctions.R:221-520, getR2 :1366-1370) is the independent check."""
    ya, yb = synth.thetalogistic(260, 71, theta=3.0), 3.0 + 2.0 * synth.thetalogistic(260, 72, theta=2.0)
    y = np.concatenate([ya, yb])
    pop = np.array(["A"] * 260 + ["B"] * 260, dtype=object)
    cells = G.fit_grid(y, [2, 3], [1, 2], pop=pop, ngpus=1, scaling="local")
    assert all(c["status"] == 0 and c["N"] > 256 for c in cells)   # two or more tiles: the lockstep path
    for c in cells:
        m = G.GPModel.from_series(y, c["E"], c["tau"], pop=pop, scaling="local")
        one = m.fit_rprop(batch.default_initparst(c["E"], False))
        st = m.fit_stats()
        m.close()
        assert one["iter"] == c["iter"]
        assert np.max(np.abs(c["pars"] - one["pars"]) / np.maximum(1, np.abs(one["pars"]))) <= 1e-9
        for k in ("R2_insample", "R2_loo", "rmse_insample", "rmse_loo"):
            assert c[k] == pytest.approx(st[k], rel=1e-9), k
    c = cells[0]
    o = O.fitGP(data=dict(y=y, pop=pop), y="y", pop="pop", E=c["E"], tau=c["tau"], scaling="local", predictmethod="loo")
    assert o["iter"] == c["iter"]
    assert c["R2_insample"] == pytest.approx(o["insampfitstats"]["R2"], rel=1e-8)
    assert c["R2_loo"] == pytest.approx(o["outsampfitstats"]["R2"], rel=1e-8)


def test_getrhomatrix_vs_oracle(gpu_lib, thetalog2pop):
    """getrhomatrix: all pair fits in one gpedm_fit_batch call (one-CTA-per-fit kernel for pairs with <= 128
    rows) against the oracle's loop of two-population fitGP calls (R/rhomatrix.R:169-182), plus the K1 pin."""
    R, names = G.getrhomatrix(thetalog2pop, y="Abundance", pop="Population", E=3, tau=1, scaling="local")
    assert names == ["PopA", "PopB"] and round(R[0, 1], 6) == 0.327254 and R[0, 1] == R[1, 0]   # README.md:86
    rng = np.random.default_rng(12)
    ys, pops = [], []
    for k, theta in enumerate([1.0, 2.5, 2.6, 4.0]):
        ys.append(synth.thetalogistic(60 + 20 * k, 40 + k, theta=theta))   # pairs of 120..200 rows: both batch paths
        pops += ["site%d" % k] * (60 + 20 * k)
    data = dict(y=np.concatenate(ys), pop=np.asarray(pops, dtype=object))
    Ro, no = O.getrhomatrix(data, y="y", pop="pop", E=2, tau=1, scaling="local", fixedpars=[np.nan, np.nan, 0.01, np.nan])
    Rg, ng = G.getrhomatrix(data, y="y", pop="pop", E=2, tau=1, scaling="local", fixedpars=[np.nan, np.nan, 0.01, np.nan])
    assert ng == no and np.array_equal(Rg, Rg.T) and np.all(np.diag(Rg) == 1.0)
    assert np.max(np.abs(Rg - Ro)) <= TOL_MODE
    # and the matrix plugs back into fitGP as the reference intends (vignettes/hierarchical.Rmd:180-190)
    g = G.fitGP(data, y="y", pop="pop", E=2, tau=1, scaling="local", rhomatrix=Rg, rhomatrix_names=ng)
    o = O.fitGP(data, y="y", pop="pop", E=2, tau=1, scaling="local", rhomatrix=Ro, rhomatrix_names=no)
    assert g["iter"] == o["iter"]
    assert np.max(np.abs(g["pars"] - o["pars"]) / np.maximum(1, np.abs(o["pars"]))) <= TOL_MODE
    g.close()


@pytest.mark.parametrize("N,M,d,npop,use_rhomat", [(130, 1, 2, 1, False), (257, 129, 3, 2, False), (500, 260, 4, 3, True),
                                                   (700, 300, 5, 3, False)])
def test_predict_new_parity_shapes(gpu_lib, N, M, d, npop, use_rhomat):
    """predict.GP newdata block + getGPgrad (R/GPfunctions.R:1012-1036, :1863-1865) at sizes straddling the
    128-tiles, with several populations and a pairwise rho matrix: mean and GP gradient against the FP64
    oracle, variance against the extended-precision oracle (the reference's own formula loses digits)."""
    Y, X, pop, parst = make_case(100 + N, N, d, npop)
    rng = np.random.default_rng(N)
    Xn = rng.standard_normal((M, d))
    popn = rng.integers(0, npop, M)
    R = names = None
    if use_rhomat:
        R = np.array([[1, .3, .6], [.3, 1, .2], [.6, .2, 1.]])
        names = [0, 1, 2]
    ref = oracle_eval(parst, Y, X, pop, rhomatrix=R, names=names)
    phi, ve, s2, rho = ref["pars"][:d], ref["pars"][d], ref["pars"][d + 1], ref["pars"][d + 2]
    om, ov, og = O.predict_newdata_core(phi, s2, rho, ve, X, Y, pop, ref["covm"]["iKVs"], Xn, popn, R, names, d > 1)
    tm, tv = O.predict_newdata_ld(phi, s2, rho, ve, X, Y, pop, Xn, popn, R, names)
    m = G.GPModel(Y, X, pop, R, names)
    m.likegrad(parst, 1, None, None, returngradonly=False)
    gm, gv, gg = m.predict_new(Xn, popn, d > 1)
    assert np.max(np.abs(gm - om)) <= TOL_MEAN
    fsd_t, fsd_g = np.sqrt(np.asarray(tv, dtype=float)), np.sqrt(gv)
    assert np.max(np.abs(fsd_g - fsd_t) / fsd_t) <= TOL_FSD
    if d > 1:
        assert np.max(np.abs(gg - og)) <= 1e-8 * max(1.0, np.max(np.abs(og)))
    m.close()
