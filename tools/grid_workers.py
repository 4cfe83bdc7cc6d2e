"""gpedm_fit_grid on one GPU with W concurrent worker threads (GPEDM_BATCH_WORKERS_PER_GPU): do two
interleaved fits fill each other's chain-bound stretches at N ~ 8190?   python tools/grid_workers.py [ncells]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpedm_b200 as G
from gpedm_b200 import synth
y = synth.ricker(8192, 1004)
Es = [int(a) for a in sys.argv[1:]] or [7, 8, 9, 10]
for rep in range(2):
    t = time.perf_counter()
    res = G.fit_grid(y, Es, [int(t) for t in os.environ.get('TAUS', '1,2').split(',')], ngpus=1)
    wall = time.perf_counter() - t
    ev = sum(r["nevals"] for r in res)
    print(json.dumps(dict(workers=os.environ.get("GPEDM_BATCH_WORKERS_PER_GPU", "default"), rep=rep, cells=len(res), wall_s=round(wall, 3),
                          evals=int(ev), ms_per_eval=round(1e3 * wall / ev, 3), ok=all(r["status"] == 0 for r in res))))
