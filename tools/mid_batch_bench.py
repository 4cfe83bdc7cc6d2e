"""Batches of mid-size fits (129 <= N <= ~1000: too big for the one-CTA kernel, too small to fill a B200):
throughput vs concurrent workers per GPU.   GPEDM_BATCH_WORKERS_PER_GPU=W python tools/mid_batch_bench.py [series_len] [nfits]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gpedm_b200 as G
from gpedm_b200 import batch, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 330
nf = int(sys.argv[2]) if len(sys.argv) > 2 else 64
probs = []
for k in range(nf):
    p, _ = batch.grid_problems(synth.ricker(n, 100 + k), [1 + k % 4], [1 + (k // 4) % 2])
    probs += p
for rep in range(2):
    t = time.perf_counter()
    res = G.fit_batch(probs, ngpus=1, want_loo=True)
    wall = time.perf_counter() - t
ev = sum(r["nevals"] for r in res)
print(json.dumps(dict(series=n, fits=len(probs), workers=os.environ.get("GPEDM_BATCH_WORKERS_PER_GPU", "default"), wall_s=round(wall, 3),
                      evals=int(ev), evals_per_s=round(ev / wall), ok=all(r["status"] == 0 for r in res))))
