"""Throughput of batches of small independent fits (N=98, d=2, two populations): lockstep one-CTA-per-fit
path vs the tiled per-handle path (GPEDM_NO_SMALL_BATCH=1)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gpedm_b200 as G
from gpedm_b200 import batch, synth
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 592
probs = []
for k in range(nf):
    ya, yb = synth.thetalogistic(51, 100 + 2 * k, theta=3.5), synth.thetalogistic(51, 101 + 2 * k, theta=2.5)
    Ya, Xa = synth.embed(ya, 2, 1)
    Yb, Xb = synth.embed(yb, 2, 1)
    probs.append(dict(Y=np.concatenate([Ya, Yb]), X=np.vstack([Xa, Xb]), pop=np.array([0] * len(Ya) + [1] * len(Yb)),
                      initparst=batch.default_initparst(2, single_pop=False)))
G.fit_batch(probs[:4], ngpus=1)
for rep in range(2):
    t = time.perf_counter()
    res = G.fit_batch(probs, ngpus=1, want_loo=True)
    dt = time.perf_counter() - t
    ne = sum(r["nevals"] for r in res)
    print(json.dumps(dict(mode="tiled" if os.environ.get("GPEDM_NO_SMALL_BATCH") else "small-batch", fits=nf, wall_s=round(dt, 4),
                          evals=int(ne), evals_per_s=round(ne / dt, 1), fits_per_s=round(nf / dt, 1),
                          max_iter=max(r["iter"] for r in res), ok=all(r["status"] == 0 for r in res))))
