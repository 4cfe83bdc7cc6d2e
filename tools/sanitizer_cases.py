"""Small end-to-end coverage for compute-sanitizer runs (memcheck / racecheck): every kernel family once at
small sizes.   compute-sanitizer --tool memcheck python tools/sanitizer_cases.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gpedm_b200 as G
from gpedm_b200 import synth, batch, _lib
quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
Y, X, pop = synth.hierarchical_thetalog(npop=3, n=(40 if quick else 110), E=3, seed0=5)
N = len(Y)
m = G.GPModel(Y, X, pop)
p0 = batch.default_initparst(3, single_pop=False)
r = m.likegrad(p0, 1.0, None, None, returngradonly=False)
print("N", N, "nll", r["nllpost"])
if not quick:
    r = m.fit_rprop(p0, options=dict(maxiter=5))
t = np.concatenate([np.arange(np.sum(pop == k)) for k in range(3)]).astype(float)
m.predict_new(X[:37] + 0.01, pop[:37], True)
m.predict_loo(0, t, None); m.predict_loo(2, t, None, True); m.predict_lto(t, 1, None, True); m.predict_sequential(t)
m.matrix(_lib.MAT_IKVS); m.matrix(_lib.MAT_CD); m.vector(_lib.VEC_LOO_VAR); m.fit_stats()
m.close()
G.getcovinv(np.eye(150) * 2 + 0.1)
y = synth.ricker(140 if quick else 400, 3)
ms = G.GPModel.from_series(y, 3, 2); ms.likegrad(batch.default_initparst(3, True), 1.0, None, None, False); ms.fit_stats(); ms.close()
probs, _ = batch.grid_problems(synth.ricker(100, 4), [1, 2], [1])
G.fit_batch(probs, ngpus=1, want_loo=True)
if not quick:
    G.fit_grid(y, [1, 2], [1], ngpus=1)
print("sanitizer cases done")
