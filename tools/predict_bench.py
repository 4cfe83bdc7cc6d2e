"""Prediction throughput on a resident fit: newdata (M points, with / without GP input-gradient),
closed-form LOO, block leave-out (lto) and sequential, at N = series-E rows.   python tools/predict_bench.py [series] [M]"""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gpedm_b200 import _lib
if os.environ.get('GPEDM_LIB'):  # A/B against another build of the library
    _lib.LIB_PATH = os.environ['GPEDM_LIB']
import gpedm_b200 as G
from gpedm_b200 import synth
from gpedm_b200.batch import default_initparst
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
M = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
E = 10
y = synth.ricker(n + M, 1004)
Y, X = synth.embed(y[:n], E, 1)
Yn, Xn = synth.embed(y, E, 1)
Xn = np.asfortranarray(Xn[-M:])
m = G.GPModel(Y, X)
N = len(Y)
m.likegrad(default_initparst(E, True), 1.0, None, None, returngradonly=False)
def timed(f, reps=3):
    f(); t = time.perf_counter()
    for _ in range(reps): f()
    return (time.perf_counter() - t) / reps
pop = None
t_new = timed(lambda: m.predict_new(Xn, pop, False))
t_grad = timed(lambda: m.predict_new(Xn, pop, True))
tt = np.arange(N, dtype=float)
t_loo = timed(lambda: m.predict_loo(0, tt, None))
t_loo3 = timed(lambda: m.predict_loo(3, tt, None))
t_lto = timed(lambda: m.predict_lto(tt, 1, None))
t_seq = timed(lambda: m.predict_sequential(tt), reps=1)
flop_new = 1.0 * M * N * N  # triangular apply U = L^-1 Cs^T (N^2 M) dominates; assembly M N exp
print(json.dumps(dict(N=N, M=M, predict_new_ms=round(1e3 * t_new, 3), predict_new_with_gpgrad_ms=round(1e3 * t_grad, 3),
                      predict_new_tflops=round(flop_new / t_new * 1e-12, 2), points_per_s=round(M / t_new),
                      loo_closed_form_ms=round(1e3 * t_loo, 3), loo_exclradius3_ms=round(1e3 * t_loo3, 3),
                      lto_exclradius1_ms=round(1e3 * t_lto, 3), sequential_ms=round(1e3 * t_seq, 2),
                      note="host wall time per call through the Python binding, inputs in host memory (H2D of Xnew and D2H of the results included)")))
