"""Run a couple of gradient-only evaluations at a given N (for ncu launch lists / captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gpedm_b200 as G
from gpedm_b200 import synth
from gpedm_b200.batch import default_initparst
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
E = int(sys.argv[3]) if len(sys.argv) > 3 else 10
Y, X = synth.embed(synth.ricker(n + E, 1004), E, 1)
m = G.GPModel(Y, X)
p0 = default_initparst(E, True)
for _ in range(reps):
    r = m.likegrad(p0, 1.0, None, None, True)
print(len(Y), r['nllpost'], m.phase_ms(), m.launch_count())
