import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gpedm_b200 as G
from gpedm_b200 import synth
from gpedm_b200.batch import default_initparst
for n in (2048, 8192):
    Y, X = synth.embed(synth.ricker(n, 1004), 10, 1)
    G.GPModel(Y, X).close()
    t = time.perf_counter()
    for _ in range(5):
        m = G.GPModel(Y, X)
        m.close()
    print(n, 'create+destroy ms', 1e3 * (time.perf_counter() - t) / 5)
    m = G.GPModel(Y, X); p0 = default_initparst(10, True)
    t = time.perf_counter(); m.likegrad(p0, 1.0, None, None, True); print(n, 'first eval ms', 1e3 * (time.perf_counter() - t))
    t = time.perf_counter()
    for _ in range(5): m.likegrad(p0, 1.0, None, None, True)
    print(n, 'eval ms', 1e3 * (time.perf_counter() - t) / 5)
    t = time.perf_counter(); m.likegrad(p0, 1.0, None, None, False); print(n, 'first full eval ms', 1e3 * (time.perf_counter() - t))
    t = time.perf_counter(); r = m.fit_rprop(p0); dt = time.perf_counter() - t; print(n, 'fit s', dt, 'nevals', r['nevals'], 'ms/eval', 1e3 * dt / r['nevals'])
    m.close()
