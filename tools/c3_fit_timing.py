"""Where does the C3 fit spend its wall time?  handle creation vs Rprop evaluations (two repetitions)."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gpedm_b200 as G
from gpedm_b200 import synth
from gpedm_b200.batch import default_initparst
Y, X, pop = synth.hierarchical_thetalog(npop=16, n=1024, E=5, seed0=1003)
names = np.array(["pop%02d" % k for k in pop], dtype=object)
p0 = default_initparst(5, single_pop=False)
for rep in range(2):
    t0 = time.perf_counter(); m = G.GPModel(Y, X, names)
    t1 = time.perf_counter(); r = m.fit_rprop(p0)
    t2 = time.perf_counter(); ms = m.bench_evals(r["parst"], 1.0, 3)
    t3 = time.perf_counter(); m.close()
    print(json.dumps(dict(rep=rep, create_s=round(t1 - t0, 3), fit_s=round(t2 - t1, 3), nevals=r["nevals"],
                          wall_ms_per_eval=round(1e3 * (t2 - t1) / r["nevals"], 2), device_ms_per_eval_at_mode=round(ms, 2),
                          close_s=round(time.perf_counter() - t3, 3))))
