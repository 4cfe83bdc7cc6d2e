"""The 40-fit E x tau grid through ONE in-process call, gpedm_fit_grid(ngpus=G) (a worker thread per GPU and resident
fit, series uploaded once per GPU, records cross-checked by the library's ncclAllGather), twice: the first call pays
context creation and memory-pool growth on every device.   python tools/grid_inproc.py [ngpus] [series_len]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpedm_b200 as G
from gpedm_b200 import synth, _lib
ng = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
y = synth.ricker(n, 1004)
for rep in range(2):
    t = time.perf_counter()
    res = G.fit_grid(y, range(1, 11), (1, 2, 3, 4), ngpus=ng)
    wall = time.perf_counter() - t
    ev = sum(r["nevals"] for r in res)
    print(json.dumps(dict(api="gpedm_fit_grid", ngpus=ng, series=n, rep=rep, workers=os.environ.get("GPEDM_BATCH_WORKERS_PER_GPU", "default"),
                          wall_s=round(wall, 3), evals=int(ev), evals_per_s=round(ev / wall, 1), ok=all(r["status"] == 0 for r in res),
                          nccl_gather_status=int(_lib.lib().gpedm_last_gather_status()))), flush=True)
