"""Quick GPU dev loop: one parity spot-check (N=700, 3 pops) + eval-rate at N=4086/8182 (+ optional sizes)."""
import json, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gpedm_b200 as G
from gpedm_b200 import synth
from gpedm_b200.batch import default_initparst
from oracle import gpedm_oracle as O

rng = np.random.default_rng(0)
N, d, npop = 700, 5, 3
X = rng.standard_normal((N, d)); pop = rng.integers(0, npop, N)
Y = np.sin(X[:, 0]) + 0.1 * rng.standard_normal(N)
parst = np.concatenate([np.log(rng.uniform(0.05, 0.6, d)), [-3.0, 0.5, 0.3]])
D = [O.distance(X[:, i]) for i in range(d)]
ref = O.getlikegrad(parst, Y, X, pop, 1, np.full(d + 2, np.nan), None, None, D, False)
m = G.GPModel(Y, X, pop)
out = m.likegrad(parst, 1, None, None, returngradonly=False)
print(json.dumps(dict(case="parity700", nll=abs(out['nllpost'] - ref['nllpost']) / abs(ref['nllpost']),
                      grad=float(np.max(np.abs(out['grad'] - ref['grad'])) / max(1, np.max(np.abs(ref['grad'])))),
                      lnL_LOO=abs(out['lnL_LOO'] - ref['lnL_LOO']) / abs(ref['lnL_LOO']), phases=m.phase_ms())))
m.close()
m = G.GPModel(Y[:94], X[:94], pop[:94]); m.likegrad(parst, 1, None, None, True); m.likegrad(parst, 1, None, None, True)
print(json.dumps(dict(case="N94", phases=m.phase_ms()))); m.close()
sizes = [int(a) for a in sys.argv[1:]] or [4096, 8192]
for n in sizes:
    y = synth.ricker(n, 1002 if n == 4096 else 1004)
    Y, X = synth.embed(y, 10, 1)
    m = G.GPModel(Y, X)
    p0 = default_initparst(10, True)
    m.bench_evals(p0, 1.0, 1)
    ms = m.bench_evals(p0, 1.0, 5)
    print(json.dumps(dict(case='bench', N=len(Y), ms_per_eval=round(ms, 3), evals_per_s=round(1000 / ms, 2), tflops=round(len(Y) ** 3 / ms * 1e-9, 2),
                          phases={k: round(v, 3) for k, v in m.phase_ms().items()}, launches=m.launch_count())))
    r = m.likegrad(p0, 1.0, None, None, True)
    print(json.dumps(dict(case='value', N=len(Y), nll=r['nllpost'], g0=r['grad'][0], gve=r['grad'][10])))
    m.close()
