import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gpedm_b200 as G
from gpedm_b200 import batch, synth, _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
y = synth.ricker(n, 1004)
probs, meta = batch.grid_problems(y, list(range(1, 11)), [1, 2, 3, 4])
per = []
for pr in probs:
    t0 = time.perf_counter()
    m = G.GPModel(pr["Y"], pr["X"])
    t1 = time.perf_counter()
    r = m.fit_rprop(batch.default_initparst(pr["X"].shape[1]))
    t2 = time.perf_counter()
    m.vector(_lib.VEC_LOO_MEAN)
    t3 = time.perf_counter()
    m.close()
    t4 = time.perf_counter()
    per.append((pr["X"].shape[1], r["nevals"], round(1e3 * (t1 - t0), 1), round(1e3 * (t2 - t1) / r["nevals"], 2), round(1e3 * (t3 - t2), 1), round(1e3 * (t4 - t3), 1)))
print("(d, nevals, create_ms, fit_ms_per_eval, vector_ms, close_ms)")
print(per)
