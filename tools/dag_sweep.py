"""Evaluation time at a few sizes (one line each); used to sweep the GPEDM_DAG_* knobs from the shell."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpedm_b200 as G
from gpedm_b200 import synth
from gpedm_b200.batch import default_initparst
tag = {k: v for k, v in os.environ.items() if k.startswith("GPEDM_")}
for n in [int(a) for a in sys.argv[1:]] or [4096, 8192]:
    Y, X = synth.embed(synth.ricker(n, 1004), 10, 1)
    m = G.GPModel(Y, X)
    p0 = default_initparst(10, True)
    m.bench_evals(p0, 1.0, 2)
    ms = m.bench_evals(p0, 1.0, 5)
    ph = m.phase_ms()
    r = m.likegrad(p0, 1.0, None, None, True)
    print(json.dumps(dict(env=tag, N=len(Y), ms=round(ms, 3), factor_ms=round(ph["potrf"] + ph["trtri"] + ph["lauum"], 3),
                          tflops=round(len(Y) ** 3 / ms * 1e-9, 2), nll=r["nllpost"].hex() if hasattr(r["nllpost"], "hex") else float(r["nllpost"]).hex(),
                          g0=float(r["grad"][0]).hex())), flush=True)
    m.close()
