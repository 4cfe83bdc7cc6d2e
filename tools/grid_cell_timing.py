"""Where does one grid cell spend its host time?  (device prep / fit / stats / destroy)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gpedm_b200 as G
from gpedm_b200 import synth, batch
y = synth.ricker(8192, 1004)
for rep in range(3):
    for E, tau in [(10, 1), (3, 2)]:
        t0 = time.perf_counter(); m = G.GPModel.from_series(y, E, tau)
        t1 = time.perf_counter(); r = m.fit_rprop(batch.default_initparst(E, True))
        t2 = time.perf_counter(); st = m.fit_stats()
        t3 = time.perf_counter(); m.close()
        t4 = time.perf_counter()
        print("E=%d tau=%d prep %.1f ms  fit %.1f ms (%d evals, %.2f ms/eval)  stats %.2f ms  close %.1f ms" %
              (E, tau, 1e3 * (t1 - t0), 1e3 * (t2 - t1), r["nevals"], 1e3 * (t2 - t1) / r["nevals"], 1e3 * (t3 - t2), 1e3 * (t4 - t3)))
        probs, meta = batch.grid_problems(y, [E], [tau])
        t0 = time.perf_counter(); res = G.fit_batch(probs, ngpus=1, want_loo=True); t1 = time.perf_counter()
        print("   host-prep fit_batch %.1f ms (%d evals, %.2f ms/eval)" % (1e3 * (t1 - t0), res[0]["nevals"], 1e3 * (t1 - t0) / res[0]["nevals"]))
