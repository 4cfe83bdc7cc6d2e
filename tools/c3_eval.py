"""Evaluation time at the C3 shape (16 pops x 1024, E=5, N=16304, rho free) with the phase split."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gpedm_b200 as G
from gpedm_b200 import synth
from gpedm_b200.batch import default_initparst
Y, X, pop = synth.hierarchical_thetalog(npop=16, n=1024, E=5, seed0=1003)
m = G.GPModel(Y, X, pop)
p0 = default_initparst(5, single_pop=False)
m.bench_evals(p0, 1.0, 1)
ms = m.bench_evals(p0, 1.0, 4)
print(json.dumps(dict(N=len(Y), ms_per_eval=round(ms, 2), phases={k: round(v, 2) for k, v in m.phase_ms().items()})))
