"""Development check on a GPU box: compares the CUDA path with the oracle on several shapes and
prints phase timings.  (The committed parity tests live in tests/; this is a quick dev loop.)"""
import json, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gpedm_b200 as G
from gpedm_b200 import _lib, synth
from gpedm_b200.batch import default_initparst
from oracle import gpedm_oracle as O

def relerr(a, b):
    a = np.asarray(a, float); b = np.asarray(b, float)
    return float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b))))

def check_case(name, Y, X, pop, parst, rhomatrix=None, names=None, fixedpars=None, rhofixed=None):
    d = X.shape[1]
    fp = np.full(d + 2, np.nan) if fixedpars is None else fixedpars
    D = [O.distance(X[:, i]) for i in range(d)]
    ref = O.getlikegrad(parst, Y, X, pop, 1, fp, rhofixed, rhomatrix, D, False, names)
    m = G.GPModel(Y, X, pop, rhomatrix, names)
    out = m.likegrad(parst, 1, fixedpars, rhofixed, returngradonly=False)
    res = dict(case=name, N=len(Y), d=d,
               nll=abs(out['nllpost'] - ref['nllpost']) / abs(ref['nllpost']),
               grad=float(np.max(np.abs(out['grad'] - ref['grad'])) / max(1, np.max(np.abs(ref['grad'])))),
               lnL_LOO=abs(out['lnL_LOO'] - ref['lnL_LOO']) / abs(ref['lnL_LOO']),
               df=abs(out['df'] - ref['df']) / abs(ref['df']),
               SS=abs(out['SS'] - ref['SS']) / abs(ref['SS']), logdet=abs(out['logdet'] - ref['logdet']) / abs(ref['logdet']))
    if len(Y) <= 1500:
        res['Sigma'] = relerr(m.matrix(_lib.MAT_SIGMA), ref['covm']['Sigma'])
        res['Cd'] = relerr(m.matrix(_lib.MAT_CD), ref['covm']['Cd'])
        res['iKVs'] = relerr(m.matrix(_lib.MAT_IKVS), ref['covm']['iKVs'])
        Lr = np.linalg.cholesky((ref['covm']['Sigma'] + ref['covm']['Sigma'].T) / 2)
        res['L'] = relerr(m.matrix(_lib.MAT_L), Lr)
        res['Linv'] = relerr(m.matrix(_lib.MAT_LINV), np.linalg.inv(Lr))
        res['mean'] = relerr(m.vector(_lib.VEC_MEAN), ref['mean'])
        res['var'] = relerr(m.vector(_lib.VEC_VAR), np.diag(ref['cov']))
    res['phase_ms'] = m.phase_ms()
    print(json.dumps(res))
    m.close()
    return res

rng = np.random.default_rng(0)
# 1-tile, 2 pops (K1-like shape)
for (N, d, npop) in [(94, 3, 2), (300, 4, 1), (700, 5, 3), (1000, 10, 1), (1500, 2, 4)]:
    X = rng.standard_normal((N, d)); pop = rng.integers(0, npop, N)
    Y = np.sin(X[:, 0]) + 0.1 * rng.standard_normal(N)
    parst = np.concatenate([np.log(rng.uniform(0.05, 0.6, d)), [-3.0, 0.5, 0.3]])
    check_case("rand", Y, X, pop, parst)
# rhomatrix + fixedpars
N, d, npop = 500, 3, 3
X = rng.standard_normal((N, d)); pop = rng.integers(0, npop, N); Y = np.cos(X[:, 1]) + 0.1 * rng.standard_normal(N)
R = np.array([[1, .3, .5], [.3, 1, .2], [.5, .2, 1.]])
check_case("rhomat", Y, X, pop, np.array([-1, -2, -1.5, -3, .4, .1]), rhomatrix=R, names=[0, 1, 2])
check_case("fixed", Y, X, pop, np.array([-1, -2, -1.5, -3, .4, .1]), fixedpars=np.array([np.nan, 0.2, np.nan, 0.01, np.nan]), rhofixed=0.3)

# full fits vs oracle
def csv(name):
    import csv as _c
    rows = list(_c.reader(open(os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden', name + '.csv'))))
    cols = rows[0]; out = {}
    for j, c in enumerate(cols):
        v = [r[j] for r in rows[1:]]
        try: out[c] = np.array([float(x) for x in v])
        except ValueError: out[c] = np.array(v, dtype=object)
    return out
tl = csv('thetalog2pop')
t = time.time()
g = G.fitGP(data=tl, y='Abundance', pop='Population', E=3, tau=1, scaling='local', predictmethod='loo')
tg = time.time() - t
o = O.fitGP(data=tl, y='Abundance', pop='Population', E=3, tau=1, scaling='local', predictmethod='loo')
print(json.dumps(dict(case='K1', secs=tg, iter=[g['iter'], o['iter']], pars_err=float(np.max(np.abs(g['pars'] - o['pars']))),
      pars=g['pars'].tolist(), R2=[g['insampfitstats']['R2'], o['insampfitstats']['R2']],
      looR2=[g['outsampfitstats']['R2'], o['outsampfitstats']['R2']],
      loo_mean_err=float(np.nanmax(np.abs(g['outsampresults']['predmean'] - o['outsampresults']['predmean']))),
      loo_sd_rel=float(np.nanmax(np.abs(g['outsampresults']['predfsd'] / o['outsampresults']['predfsd'] - 1))))))
for meth, kw in [('lto', {}), ('lto', dict(exclradius=1)), ('loo', dict(exclradius=2)), ('sequential', {}), ('loo', dict(returnGPgrad=True)), ('lto', dict(returnGPgrad=True))]:
    pg = G.predict(g, predictmethod=meth, **kw); po = O.predict(o, predictmethod=meth, **kw)
    r = dict(case='K1-' + meth, kw=str(kw), mean_err=float(np.nanmax(np.abs(pg['outsampresults']['predmean'] - po['outsampresults']['predmean']))),
             sd_rel=float(np.nanmax(np.abs(pg['outsampresults']['predsd'] / po['outsampresults']['predsd'] - 1))),
             fsd_rel=float(np.nanmax(np.abs(pg['outsampresults']['predfsd'] / po['outsampresults']['predfsd'] - 1))))
    if 'GPgrad' in pg: r['gpgrad_err'] = float(np.nanmax(np.abs(pg['GPgrad'] - po['GPgrad'])))
    print(json.dumps(r))
# newdata prediction with GP grad
Xn = g['inputs']['X'][:20] + 0.01; pn = g['inputs']['Pop'][:20]
mg, vg, gg = g['model'].predict_new(Xn, pn, True)
d = 3; ph = o['pars']
mo, vo, go = O.predict_newdata_core(ph[:d], ph[d+1], ph[d+2], ph[d], o['inputs']['X'], o['inputs']['Y'], o['inputs']['Pop'], o['covm']['iKVs'], Xn, pn, returnGPgrad=True)
print(json.dumps(dict(case='K1-new', mean_err=float(np.max(np.abs(mg - mo))), var_rel=float(np.max(np.abs(vg / vo - 1))), grad_err=float(np.max(np.abs(gg - go))))))

# performance: N=4086 d=10 and N=8182 d=10
for n in (4096, 8192):
    y = synth.ricker(n, 1002 if n == 4096 else 1004)
    Y, X = synth.embed(y, 10, 1)
    m = G.GPModel(Y, X)
    p0 = default_initparst(10, True)
    m.bench_evals(p0, 1.0, 1)
    ms = m.bench_evals(p0, 1.0, 3)
    print(json.dumps(dict(case='bench', N=len(Y), ms_per_eval=ms, evals_per_s=1000 / ms, tflops=len(Y) ** 3 / ms * 1e-9, phases=m.phase_ms(), launches=m.launch_count())))
    if n == 4096:
        t = time.time(); r = m.fit_rprop(p0); tf = time.time() - t
        print(json.dumps(dict(case='fit4096', secs=tf, iter=r['iter'], nevals=r['nevals'], pars=r['pars'].tolist(), nll=r['nllpost'])))
    m.close()
