"""Deterministic synthetic delay-embedding workloads, generated with numpy
PCG64 so the GPU path and the CPU baseline see identical inputs.  The generators mirror the
reference's simulated datasets (data-raw/datasims.R:5-13 theta-logistic, :39-53 Hastings-Powell,
:146-154 Ricker); they build inputs only and are not part of the numerical hot path."""
import numpy as np


def ricker(n, seed, r=2.8, s=0.05, burn=200):
    rng = np.random.default_rng(seed)
    y = np.empty(n + burn)
    y[0] = rng.uniform()
    eps = rng.standard_normal(n + burn)
    for i in range(1, n + burn):
        y[i] = y[i - 1] * np.exp(r * (1 - y[i - 1])) * np.exp(s * eps[i])
    return y[burn:]


def thetalogistic(n, seed, r=1.2, theta=2.5, s=0.05, burn=200):
    rng = np.random.default_rng(seed)
    y = np.empty(n + burn)
    y[0] = rng.uniform()
    eps = rng.standard_normal(n + burn)
    for i in range(1, n + burn):
        y[i] = y[i - 1] * np.exp(r * (1 - y[i - 1] ** theta)) * np.exp(s * eps[i])
    return y[burn:]


def hastings_powell(n, seed, dt=0.01, sample=1.0, noise=0.05):
    a, b, c, d, m, mu = 5.0, 3.0, 0.1, 2.0, 0.4, 0.01

    def f(v):
        X, Y, Z = v
        return np.array([X * (1 - X) - a * X * Y / (1 + b * X),
                         a * X * Y / (1 + b * X) - c * Y * Z / (1 + d * Y) - m * Y,
                         c * Y * Z / (1 + d * Y) - mu * Z])

    v = np.array([0.8, 0.1, 9.0])
    steps = int(round(sample / dt))
    out = np.empty((n, 3))
    for i in range(n):
        out[i] = v
        for _ in range(steps):
            k1 = f(v)
            k2 = f(v + 0.5 * dt * k1)
            k3 = f(v + 0.5 * dt * k2)
            k4 = f(v + dt * k3)
            v = v + dt / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
    rng = np.random.default_rng(seed)
    return out + noise * rng.standard_normal(out.shape) * out.std(axis=0)[None, :]


def embed(y, E, tau=1):
    """Single-population delay embedding + global scaling + NA-row removal: what fitGP(y, E, tau)
    hands to fmingrad_Rprop (R/GPfunctions.R:292-306, 378-385, 446-451).  Returns (Y, X)."""
    y = np.asarray(y, dtype=np.float64)
    n = len(y)
    X = np.full((n, E), np.nan)
    for i in range(1, E + 1):
        X[tau * i:, i - 1] = y[: n - tau * i]
    ys = (y - y.mean()) / y.std(ddof=1)
    xm = np.nanmean(X, axis=0)
    xs = np.nanstd(X, axis=0, ddof=1)
    Xs = (X - xm[None, :]) / xs[None, :]
    ok = ~np.isnan(Xs).any(axis=1)
    return ys[ok], Xs[ok]


def hierarchical_thetalog(npop=16, n=1024, E=5, seed0=1003):
    """Config 3: npop theta-logistic populations, local scaling, E lags; returns (Y, X, pop)."""
    Ys, Xs, pops = [], [], []
    for k in range(npop):
        y = thetalogistic(n, seed0 + k, theta=2.5 + k / 15.0)
        Yk, Xk = embed(y, E, 1)
        Ys.append(Yk)
        Xs.append(Xk)
        pops.append(np.full(len(Yk), k, dtype=np.int64))
    return np.concatenate(Ys), np.vstack(Xs), np.concatenate(pops)


def hastings_powell_fast(n, seed, dt=0.1, sample=1.0, noise=0.05):
    """Same system and sampling as `hastings_powell` with the RK4 steps in plain Python floats (10x faster
    than 3-element numpy arrays; 32768 samples in a few seconds at dt=0.1)."""
    a, b, c, d, m, mu = 5.0, 3.0, 0.1, 2.0, 0.4, 0.01

    def f(X, Y, Z):
        g = a * X * Y / (1 + b * X)
        h = c * Y * Z / (1 + d * Y)
        return X * (1 - X) - g, g - h - m * Y, h - mu * Z

    X, Y, Z = 0.8, 0.1, 9.0
    steps = int(round(sample / dt))
    out = np.empty((n, 3))
    for i in range(n):
        out[i] = (X, Y, Z)
        for _ in range(steps):
            k1 = f(X, Y, Z)
            k2 = f(X + 0.5 * dt * k1[0], Y + 0.5 * dt * k1[1], Z + 0.5 * dt * k1[2])
            k3 = f(X + 0.5 * dt * k2[0], Y + 0.5 * dt * k2[1], Z + 0.5 * dt * k2[2])
            k4 = f(X + dt * k3[0], Y + dt * k3[1], Z + dt * k3[2])
            X += dt / 6 * (k1[0] + 2 * k2[0] + 2 * k3[0] + k4[0])
            Y += dt / 6 * (k1[1] + 2 * k2[1] + 2 * k3[1] + k4[1])
            Z += dt / 6 * (k1[2] + 2 * k2[2] + 2 * k3[2] + k4[2])
    rng = np.random.default_rng(seed)
    return out + noise * rng.standard_normal(out.shape) * out.std(axis=0)[None, :]


def hastings_powell_embedding(N, seed=1005, lags=3):
    """Config 5: y = species X, predictors = `lags` lags each of X, Y, Z (d = 3 * lags), global scaling,
    exactly N complete rows (series length N + lags).  Returns (Y, X)."""
    hp = hastings_powell_fast(N + lags, seed)
    n = len(hp)
    cols = []
    for sp in range(3):
        for i in range(1, lags + 1):
            c = np.full(n, np.nan)
            c[i:] = hp[: n - i, sp]
            cols.append(c)
    Xr = np.column_stack(cols)
    y = hp[:, 0]
    ys = (y - y.mean()) / y.std(ddof=1)
    Xs = (Xr - np.nanmean(Xr, axis=0)[None, :]) / np.nanstd(Xr, axis=0, ddof=1)[None, :]
    ok = ~np.isnan(Xs).any(axis=1)
    return ys[ok], Xs[ok]
