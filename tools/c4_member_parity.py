"""Full-size parity of the headline workload's FIT: the C4 grid member (series 8192, E=10, tau=1 -> N=8182, d=10)
fitted by the CPU oracle (`oracle` mode: ~107 evaluations of 10-20 s each, run in the build container) and
by the GPU library (`gpu` mode, on a B200), then compared (`compare`).
  python tools/c4_member_parity.py oracle|gpu|compare"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from gpedm_b200 import synth

mode = sys.argv[1]
y = synth.ricker(8192, 1004)
keys = ("R2", "rmse", "ln_post", "ln_prior", "lnL_LOO", "df", "SS", "logdet")
if mode == "oracle":
    from oracle import gpedm_oracle as O
    t = time.perf_counter()
    o = O.fitGP(y=y, E=10, tau=1)
    rec = dict(secs=time.perf_counter() - t, iter=int(o["iter"]), pars=[float(v) for v in o["pars"]],
               stats={k: float(o["insampfitstats"][k]) for k in keys},
               predmean=[float(v) for v in o["insampresults"]["predmean"][::64]],
               predsd=[float(v) for v in o["insampresults"]["predsd"][::64]])
    json.dump(rec, open(os.path.join(ROOT, "gpurun_out", "c4_oracle.json"), "w"))
    print("oracle done", rec["secs"], rec["iter"])
elif mode == "gpu":
    import gpedm_b200 as G
    t = time.perf_counter()
    g = G.fitGP(y=y, E=10, tau=1)
    rec = dict(secs=time.perf_counter() - t, iter=int(g["iter"]), pars=[float(v) for v in g["pars"]],
               stats={k: float(g["insampfitstats"][k]) for k in keys},
               predmean=[float(v) for v in g["insampresults"]["predmean"][::64]],
               predsd=[float(v) for v in g["insampresults"]["predsd"][::64]])
    json.dump(rec, open(os.path.join(ROOT, "gpurun_out", "c4_gpu.json"), "w"))
    print("gpu done", rec["secs"], rec["iter"])
else:
    o = json.load(open(os.path.join(ROOT, "gpurun_out", "c4_oracle.json")))
    g = json.load(open(os.path.join(ROOT, "gpurun_out", "c4_gpu.json")))
    po, pg = np.array(o["pars"]), np.array(g["pars"])
    rec = dict(config="C4 member (series 8192, E=10, tau=1, N=8182): full fit, GPU vs CPU oracle (oracle run in the build container)",
               oracle_fit_s=o["secs"], gpu_fit_s=g["secs"], iter=[g["iter"], o["iter"]],
               max_mode_err=float(np.max(np.abs(pg - po) / np.maximum(1, np.abs(po)))),
               stats_rel={k: abs(g["stats"][k] - o["stats"][k]) / max(1e-300, abs(o["stats"][k])) for k in keys},
               predmean_max_abs_err=float(np.nanmax(np.abs(np.array(g["predmean"]) - np.array(o["predmean"])))),
               predsd_max_rel_err=float(np.nanmax(np.abs(np.array(g["predsd"]) / np.array(o["predsd"]) - 1))))
    json.dump(rec, open(os.path.join(ROOT, "profiles", "r01_c4_member_parity_vs_oracle.json"), "w"), indent=1)
    print(json.dumps(rec, indent=1))
