"""Time the 40-fit E x tau grid on one GPU (two consecutive calls: the first includes CUDA context
creation and lazy kernel loading)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gpedm_b200 as G
from gpedm_b200 import batch, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
y = synth.ricker(n, 1004)
probs, meta = batch.grid_problems(y, list(range(1, 11)), [1, 2, 3, 4])
for rep in range(reps):
    t = time.perf_counter()
    res = G.fit_batch(probs, ngpus=1, want_loo=True)
    wall = time.perf_counter() - t
    print(json.dumps(dict(n=n, rep=rep, workers=os.environ.get("GPEDM_BATCH_WORKERS_PER_GPU", "2"), wall_s=round(wall, 3),
                          evals=int(sum(r["nevals"] for r in res)), ms_per_eval=round(1e3 * wall / sum(r["nevals"] for r in res), 3),
                          ok=all(r["status"] == 0 for r in res))))
