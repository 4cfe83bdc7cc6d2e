"""Full-size GPU parity (VERDICT r01 row g): the BASELINE.json configurations C2-C5 at their own sizes,
through the C ABI, against the CPU oracle where it finishes in about a minute and against independent
host LAPACK / brute-force refactorisations where it does not.  Also the variance outputs of every
prediction method against the extended-precision (80-bit) oracle at the north-star 1e-9.

Tolerances: nll rel 1e-10, gradient 1e-8 (max abs / max(1, |g|_inf)), predmean 1e-9,
predfsd rel 1e-9 vs the 80-bit oracle."""
import time

import numpy as np
import pytest

import gpedm_b200 as G
from gpedm_b200 import _lib, batch, synth
from oracle import gpedm_oracle as O

pytestmark = pytest.mark.gpu

TOL_NLL, TOL_GRAD, TOL_MEAN, TOL_FSD = 1e-10, 1e-8, 1e-9, 1e-9


def grad_err(g, ref):
    return float(np.max(np.abs(g - ref)) / max(1.0, np.max(np.abs(ref))))


def rel(a, b):
    return abs(a - b) / abs(b)


def oracle_gradonly(parst, Y, X, pop):
    """oracle.getlikegrad with returngradonly=TRUE (R/GPfunctions.R:586-725): no N^3 dgemm extras."""
    d = X.shape[1]
    D = [O.distance(X[:, i]) for i in range(d)]
    return O.getlikegrad(parst, Y, X, pop, 1, np.full(d + 2, np.nan), None, None, D, True)


def mixed_start(d, single_pop=True):
    """A start with distinct length scales (so every phi trace term differs) and a small ve."""
    p = batch.default_initparst(d, single_pop)
    p[:d] = np.log(np.array([0.3, 0.1, 0.05, 0.02, 0.01, 0.01, 0.02, 0.01, 0.03, 0.01])[:d])
    p[d] = -4.0
    return p


# ---------------------------------------------------------------------------- C4 member, headline path
def test_C4_member_N8182_nll_and_gradient_vs_oracle(gpu_lib):
    """The bench.py workload itself (series 8192, E=10, tau=1 -> N=8182, nb=64: live three-stream schedule,
    grouped rank-512 updates), one gradient-only evaluation: nll AND every gradient component against the
    oracle, at the default start and at a start with distinct length scales."""
    Y, X = synth.embed(synth.ricker(8192, 1004), 10, 1)
    pop = np.ones(len(Y), dtype=int)
    m = G.GPModel(Y, X)
    for p0 in (batch.default_initparst(10, True), mixed_start(10)):
        ref = oracle_gradonly(p0, Y, X, pop)
        out = m.likegrad(p0, 1, None, None, True)
        assert rel(out["nllpost"], ref["nllpost"]) <= TOL_NLL
        assert grad_err(out["grad"], ref["grad"]) <= TOL_GRAD
    m.close()


def test_live_stream_group2_path_N6500_vs_oracle(gpu_lib):
    """49 <= nb < 64: above the CUDA-graph limit (live streams) but with group size 2."""
    Y, X, pop = synth.hierarchical_thetalog(npop=2, n=3253, E=3, seed0=300)
    N = len(Y)
    assert 48 * 128 < N <= 63 * 128
    p0 = np.concatenate([np.log([0.4, 0.1, 0.03]), [-3.0, 0.5, 0.4]])
    ref = oracle_gradonly(p0, Y, X, pop)
    m = G.GPModel(Y, X, pop)
    out = m.likegrad(p0, 1, None, None, True)
    assert rel(out["nllpost"], ref["nllpost"]) <= TOL_NLL
    assert grad_err(out["grad"], ref["grad"]) <= TOL_GRAD
    m.close()


# ---------------------------------------------------------------------------- C2
def test_C2_N4086_full_evaluation_and_predict_vs_oracle(gpu_lib):
    """C2: Ricker series 4096, E=10, tau=1, global scaling -> N=4086: one FULL evaluation (in-sample mean,
    lnL_LOO, df) and the 1024-point newdata prediction against the oracle."""
    y = synth.ricker(4096 + 1024, 1002)
    Y, X = synth.embed(y[:4096], 10, 1)
    Yn, Xn = synth.embed(y[4096 - 10:], 10, 1)
    Xn = Xn[:1024]
    N, d = len(Y), 10
    assert N == 4086
    pop = np.ones(N, dtype=int)
    p0 = mixed_start(10)
    D = [O.distance(X[:, i]) for i in range(d)]
    ref = O.getlikegrad(p0, Y, X, pop, 1, np.full(d + 2, np.nan), None, None, D, False)
    m = G.GPModel(Y, X)
    out = m.likegrad(p0, 1, None, None, False)
    assert rel(out["nllpost"], ref["nllpost"]) <= TOL_NLL
    assert grad_err(out["grad"], ref["grad"]) <= TOL_GRAD
    assert rel(out["lnL_LOO"], ref["lnL_LOO"]) <= 1e-9 and rel(out["df"], ref["df"]) <= 1e-9
    np.testing.assert_allclose(m.vector(_lib.VEC_MEAN), ref["mean"], atol=TOL_MEAN)
    gm, gv, _ = m.predict_new(Xn, np.ones(1024, dtype=int))
    p = ref["pars"]
    _, Cs = O.getcov(p[:d], p[d + 1], p[d + 2], X, Xn, pop, np.ones(1024, dtype=int))
    om = Cs @ ref["a"]
    assert np.max(np.abs(gm - om)) <= TOL_MEAN
    # variance: sigma2 - ||L^-1 c||^2 with a host LAPACK factor of the oracle's Sigma (FP64; 1e-7 is what two
    # FP64 evaluations of this cancellation-prone quantity agree to - the 1e-9 claim is tested vs 80-bit below)
    from scipy.linalg import cholesky, solve_triangular
    L = cholesky(ref["covm"]["Sigma"], lower=True)
    U = solve_triangular(L, Cs.T, lower=True)
    ov = p[d + 1] - np.sum(U * U, axis=0)
    assert np.max(np.abs(np.sqrt(gv) / np.sqrt(ov) - 1)) <= 1e-7
    m.close()


# ---------------------------------------------------------------------------- C3 shape
def test_C3_shape_16pops_free_rho_vs_oracle_and_bruteforce_loo(gpu_lib):
    """C3's shape - 16 populations, free rho (no rhomatrix), E=5, local scaling - at 16 x 640 points
    (N=10160, nb=80; the full C3 N=16304 takes the oracle several minutes): evaluation vs the oracle, and
    the closed-form LOO of 8 random rows vs brute-force refactorisations of the (N-1)x(N-1) matrix
    (reference R/GPfunctions.R:1056-1086)."""
    Y, X, pop = synth.hierarchical_thetalog(npop=16, n=640, E=5, seed0=1003)
    N, d = len(Y), 5
    assert N == 16 * 635
    p0 = np.concatenate([np.log([0.5, 0.1, 0.02, 0.01, 0.01]), [-3.5, 0.4, -0.3]])
    D = [O.distance(X[:, i]) for i in range(d)]
    ref = O.getlikegrad(p0, Y, X, pop, 1, np.full(d + 2, np.nan), None, None, D, True)
    del D
    m = G.GPModel(Y, X, pop)
    out = m.likegrad(p0, 1, None, None, False)
    assert rel(out["nllpost"], ref["nllpost"]) <= TOL_NLL
    assert grad_err(out["grad"], ref["grad"]) <= TOL_GRAD
    assert abs(out["grad"][d + 2]) > 1e-3  # the rho trace term is live
    lm, lv, _ = m.predict_loo(0, None, None)
    p = out["pars"]
    _, Cd = O.getcov(p[:d], p[d + 1], p[d + 2], X, X, pop, pop)
    Sigma = Cd + p[d] * np.eye(N)
    from scipy.linalg import cho_factor, cho_solve
    rng = np.random.default_rng(7)
    for i in rng.choice(N, 8, replace=False):
        keep = np.delete(np.arange(N), i)
        cf = cho_factor(Sigma[np.ix_(keep, keep)], lower=True, overwrite_a=True, check_finite=False)
        c = Cd[i, keep]
        w = cho_solve(cf, c, check_finite=False)
        mean_i = w @ Y[keep]
        var_i = p[d + 1] - c @ w
        assert abs(lm[i] - mean_i) <= TOL_MEAN
        assert abs(np.sqrt(lv[i] + p[d]) / np.sqrt(var_i + p[d]) - 1) <= 1e-7  # FP64 brute force: see 80-bit test
    m.close()


# ---------------------------------------------------------------------------- C5
def test_C5_N32768_against_host_lapack(gpu_lib):
    """C5: N=32768, d=9 (3 lags of 3 species), single large factorisation.  The oracle's 30-odd N x N
    temporaries do not fit a unit test at this size, so: (1) rows of the GPU's Sigma against the oracle's
    getcov entry by entry; (2) log-determinant, SS = Y'a and `like` against a host LAPACK dpotrf of that
    Sigma; (3) || Sigma a - Y ||_inf; (4) df = N - ve tr(Sigma^-1) and lnL_LOO against the same factor via
    a = Sigma^-1 Y identities on a random subset of the inverse diagonal."""
    N, d = 32768, 9
    Y, X = synth.hastings_powell_embedding(N, seed=1005)
    assert X.shape == (N, d)
    pop = np.ones(N, dtype=int)
    p0 = np.concatenate([np.log([0.3, 0.05, 0.02, 0.2, 0.03, 0.01, 0.1, 0.02, 0.01]), [-3.0, 0.3, 0.0]])
    m = G.GPModel(Y, X)
    t0 = time.time()
    out = m.likegrad(p0, 1, None, None, False)
    t_gpu = time.time() - t0
    a = m.vector(_lib.VEC_ALPHA)
    diag_s = m.vector(_lib.VEC_DIAG_IKVS)
    Sigma = m.matrix(_lib.MAT_SIGMA)
    m.close()
    p = out["pars"]
    phi, ve, s2 = p[:d], p[d], p[d + 1]
    rows = np.array([0, 1, 127, 128, 129, 8191, 16384, 20000, 32767])
    _, Cr = O.getcov(phi, s2, 0.5, X, X[rows], pop, pop[rows])
    Cr[np.arange(len(rows)), rows] += ve
    np.testing.assert_allclose(Sigma[rows], Cr, rtol=4e-15, atol=1e-300)
    np.testing.assert_array_equal(Sigma[rows], Sigma[:, rows].T)
    resid = Sigma @ a - Y
    assert np.max(np.abs(resid)) <= 5e-7 * max(1.0, np.max(np.abs(a)) * ve)
    from scipy.linalg import lapack, solve_triangular
    t0 = time.time()
    L, info = lapack.dpotrf(Sigma, lower=1, overwrite_a=1)
    t_cpu = time.time() - t0
    assert info == 0
    logdet = np.log(np.diag(L)).sum()
    z = solve_triangular(L, Y, lower=True, check_finite=False)
    assert rel(out["logdet"], logdet) <= 1e-10
    assert rel(out["SS"], z @ z) <= 1e-9
    like = -0.5 * (z @ z) - logdet
    assert rel(out["like"], like) <= TOL_NLL
    ah = solve_triangular(L, z, lower=True, trans="T", check_finite=False)
    np.testing.assert_allclose(a, ah, rtol=1e-6, atol=1e-8 * np.max(np.abs(ah)))
    # inverse diagonal on a random subset: S_ii = || L^-1 e_i ||^2
    idx = np.sort(np.random.default_rng(1).choice(N, 16, replace=False))
    E = np.zeros((N, len(idx)))
    E[idx, np.arange(len(idx))] = 1.0
    V = solve_triangular(L, E, lower=True, check_finite=False)
    np.testing.assert_allclose(diag_s[idx], np.sum(V * V, axis=0), rtol=1e-8)
    assert abs(out["df"] - (N - ve * diag_s.sum())) <= 1e-6 * N
    print("C5 full evaluation on the GPU %.2f s; host dpotrf alone %.1f s" % (t_gpu, t_cpu))


# ---------------------------------------------------------------------------- variances vs 80-bit
@pytest.fixture(scope="module")
def k1(gpu_lib, thetalog2pop):
    g = G.fitGP(data=thetalog2pop, y="Abundance", pop="Population", E=3, tau=1, scaling="local", predictmethod="loo")
    yield g
    g.close()


def test_all_prediction_sds_against_extended_precision(k1):
    """predfsd of EVERY prediction method (in-sample, loo r=0 / r=2, lto r=0 / r=1, sequential) within 1e-9
    relative of the 80-bit oracle on the README model (K1: N=94, 2 populations); the FP64 restatement of
    the reference's own formulas is 1e-9..1e-6 away from it, so round 1's 1e-6/1e-7
    comparisons against it could not show this."""
    d = 3
    p = k1["pars"]
    inp = k1["inputs"]
    X, Y, Pop, Time, Prim = inp["X"], inp["Y"], inp["Pop"], inp["Time"], inp["Primary"]
    phi, ve, s2, rho = p[:d], p[d], p[d + 1], p[d + 2]
    model = k1["model"]
    worst = {}
    # in-sample: diag(Cd - Cd iKVs Cd) == ve - ve^2 Sigma^-1_ii
    _, a_t, dinv = O.loglik_ld(phi, s2, rho, ve, X, Y, Pop)
    ld = np.longdouble
    truth = ld(ve) - ld(ve) ** 2 * dinv
    got = model.vector(_lib.VEC_VAR)
    worst["insample"] = float(np.max(np.abs(np.sqrt(got) / np.sqrt(np.asarray(truth, float)) - 1)))
    mean_t = np.asarray(Y, ld) - ld(ve) * a_t
    assert np.max(np.abs(model.vector(_lib.VEC_MEAN) - np.asarray(mean_t, float))) <= TOL_MEAN
    for method, r in (("loo", 0), ("loo", 2), ("lto", 0), ("lto", 1)):
        tm, tv = O.predict_leaveout_ld(method, phi, s2, rho, ve, X, Y, Pop, Time, Prim, r)
        if method == "loo":
            gm, gv, _ = model.predict_loo(r, Time, None)
        else:
            gm, gv, _ = model.predict_lto(Time, r, None)
        assert np.max(np.abs(gm - np.asarray(tm, float))) <= TOL_MEAN, (method, r)
        worst["%s_r%d" % (method, r)] = float(np.max(np.abs(np.sqrt(gv) / np.sqrt(np.asarray(tv, float)) - 1)))
    sm, ss = O.predict_sequential_ld(phi, s2, rho, ve, X, Y, Pop, Time)
    gm, gs = model.predict_sequential(Time)
    assert np.max(np.abs(gm - np.asarray(sm, float))) <= TOL_MEAN
    worst["sequential"] = float(np.max(np.abs(gs / np.asarray(ss, float) - 1)))
    print("predfsd rel err vs 80-bit oracle:", worst)
    for k, v in worst.items():
        assert v <= TOL_FSD, (k, v)


def test_unseen_population_in_newdata(gpu_lib):
    """getcov with a new-data population absent from the training set (reference R/GPfunctions.R:796-811):
    popsame = 0 against every training row -> the rho branch; under a rhomatrix the entries keep 1."""
    rng = np.random.default_rng(14)
    X, X2 = rng.standard_normal((140, 2)), rng.standard_normal((9, 2))
    pop = np.array(["a", "b"], dtype=object)[rng.integers(0, 2, 140)]
    pop2 = np.array(["a", "zz", "b", "zz", "a", "b", "zz", "a", "b"], dtype=object)
    phi = np.array([0.4, 0.9])
    ps, Cd = O.getcov(phi, 1.3, 0.35, X, X2, pop, pop2)
    got = G.getcov(phi, 1.3, 0.35, X, X2, pop, pop2)
    np.testing.assert_allclose(got["Cd"], Cd, rtol=2e-15, atol=1e-300)
    np.testing.assert_array_equal(got["popsame"], ps)
    R = np.array([[1, .2], [.2, 1.]])
    names = O.unique_stable(pop)
    _, Cd = O.getcov(phi, 1.3, 0.35, X, X2, pop, pop2, R, names)
    got = G.getcov(phi, 1.3, 0.35, X, X2, pop, pop2, R, names)
    np.testing.assert_allclose(got["Cd"], Cd, rtol=2e-15, atol=1e-300)
    # and through predict_new on a resident fit
    Y = np.sin(X[:, 0]) + 0.1 * rng.standard_normal(140)
    parst = np.array([np.log(0.4), np.log(0.9), -3.0, 0.5, 0.3])
    m = G.GPModel(Y, X, pop)
    full = m.likegrad(parst, 1, None, None, False)
    d = 2
    D = [O.distance(X[:, i]) for i in range(d)]
    ref = O.getlikegrad(parst, Y, X, pop, 1, np.full(d + 2, np.nan), None, None, D, False)
    pr = ref["pars"]
    om, ov, _ = O.predict_newdata_core(pr[:d], pr[d + 1], pr[d + 2], pr[d], X, Y, pop, ref["covm"]["iKVs"], X2, pop2)
    gm, gv, _ = m.predict_new(X2, pop2)
    assert np.max(np.abs(gm - om)) <= TOL_MEAN
    assert np.max(np.abs(np.sqrt(gv) / np.sqrt(ov) - 1)) <= 1e-6
    m.close()
