"""Task timeline of the persistent factorisation kernel (dag_factor.cuh) from a -DGPEDM_TIMELINE build.

  nvcc ... -DGPEDM_TIMELINE -o build/libgpedm_timeline.so gpedm_b200/csrc/gpedm_core.cu -ldl
  python tools/dag_timeline.py [series_len] [bin_us]

Every task logs (globaltimer start, end, SM id, type | i << 8 | j << 20 | k tiles << 32 | wait ns << 44); wait = time
warp 0 spent polling ready flags.  Printed: SM busy / waiting fractions per time bin and task type, the chain (when
each D_c became ready), and totals."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gpedm_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'build', 'libgpedm_timeline.so')
import gpedm_b200 as G
from gpedm_b200 import synth
from gpedm_b200.batch import default_initparst

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
bin_us = float(sys.argv[2]) if len(sys.argv) > 2 else 500.0
E = 10
Y, X = synth.embed(synth.ricker(n, 1004), E, 1)
m = G.GPModel(Y, X)
p0 = default_initparst(E, True)
L = G.lib()
L.gpedm_debug_timeline.argtypes = [C.POINTER(C.c_ulonglong), C.c_uint, C.POINTER(C.c_uint)]
for _ in range(2):
    m.likegrad(p0, 1.0, None, None, True)
cnt = C.c_uint(0)
L.gpedm_debug_timeline(None, 0, C.byref(cnt))
m.likegrad(p0, 1.0, None, None, True)
MAXR = 1 << 18
buf = np.zeros(4 * MAXR, dtype=np.uint64)
L.gpedm_debug_timeline(buf.ctypes.data_as(C.POINTER(C.c_ulonglong)), MAXR, C.byref(cnt))
rec = buf[:4 * cnt.value].reshape(cnt.value, 4).astype(np.int64)
code = rec[:, 3]
marks = rec[(code & 0xff) == 150]
isd = (code & 0xff) >= 200
final_piece = ((code & 0xff) - 200) < 10
rec, code, final_piece = rec[isd], code[isd], final_piece[isd]
t0 = rec[:, 0].min()
st, en, sm = (rec[:, 0] - t0) * 1e-3, (rec[:, 1] - t0) * 1e-3, rec[:, 2]
typ, ti, tj, nkt, wait = ((code & 0xff) - 200) % 10, (code >> 8) & 0xfff, (code >> 20) & 0xfff, (code >> 32) & 0xfff, ((code >> 44) & 0xfffff) * 1e-3
T = en.max()
nsm = int(sm.max()) + 1
print("N=%d nb=%d tasks=%d SMs=%d kernel span %.3f ms; phases %s" % (len(Y), (len(Y) + 127) // 128, len(rec), nsm, T * 1e-3,
      {k: round(v, 3) for k, v in m.phase_ms().items()}))
names = ["diag", "potrf", "trtri", "lauum"]
nbin = int(np.ceil(T / bin_us))
edges = np.arange(nbin + 1) * bin_us
print("\nfraction of the %d SMs per %.0f us bin: resident per task type, of which waiting on flags (approx: wait spread uniformly over the task)" % (nsm, bin_us))
print(" ".join("%9s" % h for h in ["t_ms", "diag", "potrf", "trtri", "lauum", "waiting", "idle"]))
for b in range(nbin):
    row, wsum = [], 0.0
    for k in range(4):
        msk = typ == k
        ov = np.clip(np.minimum(en[msk], edges[b + 1]) - np.maximum(st[msk], edges[b]), 0, None)
        dur = np.maximum(en[msk] - st[msk], 1e-9)
        row.append(ov.sum() / (nsm * bin_us))
        wsum += (ov * wait[msk] / dur).sum() / (nsm * bin_us)
    print(" ".join(["%9.2f" % (edges[b] * 1e-3)] + ["%9.3f" % v for v in row] + ["%9.3f" % wsum, "%9.3f" % (1 - sum(row))]))
tot_res = (en - st).sum()
print("\ntotals (SM-ms): resident %.1f of %.1f available; waiting on flags %.1f; k tiles executed %d (%.1f us per k tile while not waiting)"
      % (tot_res * 1e-3, T * nsm * 1e-3, wait.sum() * 1e-3, nkt.sum(), (tot_res - wait.sum()) / max(1, nkt.sum() + 0.6 * ((typ > 0) & final_piece).sum())))
for k in range(4):
    msk = typ == k
    print("  %-6s %5d tasks, mean %.1f us, max %.1f us, waiting %.1f%% of residency" % (names[k], msk.sum(), (en - st)[msk].mean(), (en - st)[msk].max(),
          100 * wait[msk].sum() / max(1e-9, (en - st)[msk].sum())))
# per-task cost model: duration = a + b * (k tiles) over the tasks that never waited for a flag; b is the k loop's
# time per 128^3 k tile (17.5 us = the DMMA pipe's 2048 m8n8k4 per sub-partition at 16.05 cycles, 1.965 GHz),
# a what a task costs around its k loop (claim, flag scan, accumulator load, pipeline fill, epilogue, store, flag)
print("\ntask cost model over tasks that did not wait (least squares): us per task + us per k tile")
for k in range(4):
    msk = (typ == k) & (wait < 0.5) & final_piece & (nkt > 0)
    if msk.sum() > 8:
        A = np.stack([np.ones(msk.sum()), nkt[msk].astype(float)], axis=1)
        coef, *_ = np.linalg.lstsq(A, (en - st)[msk], rcond=None)
        resid = (en - st)[msk] - A @ coef
        print("  %-6s %5d tasks: %6.2f us + %6.3f us x k tiles (rms residual %.1f us); sum a = %.1f SM-ms, sum b k = %.1f SM-ms"
              % (names[k], msk.sum(), coef[0], coef[1], np.sqrt((resid ** 2).mean()), coef[0] * (typ == k).sum() * 1e-3, coef[1] * nkt[typ == k].sum() * 1e-3))
# phase marks (every 4th task, thread 0): 30 k loop entered, 31 k loop left, 32 resident tile in smem, 33 epilogue
# product done; matched to the task record of the same SM that encloses them
pm = marks[(((marks[:, 3] >> 8) & 0xff) >= 30) & (((marks[:, 3] >> 8) & 0xff) < 40)]
if len(pm):
    ph = {}
    by_sm = {}
    for q in range(len(rec)):
        by_sm.setdefault(int(sm[q]), []).append(q)
    for sid in by_sm:
        by_sm[sid].sort(key=lambda q: st[q])
    starts = {sid: np.array([st[q] for q in qs]) for sid, qs in by_sm.items()}
    for tm, _, sid, cd in pm:
        sid = int(sid)
        tt = (tm - t0) * 1e-3
        pos = int(np.searchsorted(starts[sid], tt, side="right")) - 1
        if pos >= 0:
            q = by_sm[sid][pos]
            if tt <= en[q] + 1e-6:
                ph.setdefault(q, {})[int((cd >> 8) & 0xff)] = tt
    print("\ntask phases (us, median over the sampled tasks that did not wait for flags; k loop per k tile):")
    print("  %-6s %6s %9s %12s %10s %10s %12s" % ("type", "tasks", "prologue", "k loop/tile", "T -> smem", "epilogue", "store+flag"))
    for k in range(4):
        rows = [(q, v) for q, v in ph.items() if typ[q] == k and wait[q] < 0.5 and final_piece[q] and nkt[q] > 0 and 30 in v and 31 in v]
        if not rows:
            continue
        pro = np.median([v[30] - st[q] for q, v in rows])
        kl = np.median([(v[31] - v[30]) / nkt[q] for q, v in rows])
        if k == 3:
            print("  %-6s %6d %9.2f %12.3f %10s %10s %12.2f" % (names[k], len(rows), pro, kl, "-", "-", np.median([en[q] - v[31] for q, v in rows])))
        elif k == 0:
            print("  %-6s %6d %9.2f %12.3f %10.2f %10s %12s" % (names[k], len(rows), pro, kl, np.median([v[32] - v[31] for q, v in rows if 32 in v]), "(chain)", "(chain)"))
        else:
            r2 = [(q, v) for q, v in rows if 32 in v and 33 in v]
            print("  %-6s %6d %9.2f %12.3f %10.2f %10.2f %12.2f" % (names[k], len(rows), pro, kl, np.median([v[32] - v[31] for q, v in r2]),
                  np.median([v[33] - v[32] for q, v in r2]), np.median([en[q] - v[33] for q, v in r2])))
d = np.where((typ == 0) & final_piece)[0]
order = d[np.argsort(tj[d])]
ends = en[order]
print("\nchain: D_c ready at (us) and step from the previous column; diag task start, its wait, its k tiles")
for q, ix in enumerate(order):
    if q % max(1, len(order) // 32) == 0 or q == len(order) - 1:
        print("  c=%3d ready %9.1f  step %7.1f  task start %9.1f wait %7.1f nkt %d" % (tj[ix], en[ix], en[ix] - (ends[q - 1] if q else 0), st[ix], wait[ix], nkt[ix]))
steps = np.diff(ends)
print("chain step: median %.1f us, mean %.1f, max %.1f; last D ready at %.1f us of %.1f" % (np.median(steps), steps.mean(), steps.max(), ends[-1], T))

# ---- chain marks (thread 0 of the two chain tasks of each block column)
if len(marks):
    mk = {}
    for tm, _, _, cd in marks[((marks[:, 3] >> 8) & 0xff) < 30]:
        mk[(int((cd >> 20) & 0xfff), int((cd >> 8) & 0xff))] = (tm - t0) * 1e-3
    names_m = [("hop: D_c flag -> seen by POTRF(c+1,c)", (5, 0), (11, 0)), ("POTRF epilogue T D^T", (11, 0), (12, 0)),
               ("POTRF store + fence + flag", (12, 0), (13, 0)), ("  flag of L(c+1,c) seen by DIAG(c+1) warp 0", (13, 0), (20, 1)), ("  first slab of the last k tile landed", (20, 1), (21, 1)),
               ("  slab 2", (21, 1), (22, 1)), ("  slab 3", (22, 1), (23, 1)), ("  slab 4", (23, 1), (24, 1)), ("  slab 5", (24, 1), (25, 1)),
               ("  slab 6", (25, 1), (26, 1)), ("  slab 7", (26, 1), (27, 1)), ("  slab 8", (27, 1), (28, 1)), ("  last slab -> k loop exit", (28, 1), (0, 1)),
               ("hop + last k tile of DIAG(c+1)", (13, 0), (0, 1)),
               ("acc -> smem", (0, 1), (1, 1)), ("Cholesky 128", (1, 1), (2, 1)), ("L -> registers, log", (2, 1), (3, 1)),
               ("inverse 128", (3, 1), (4, 1)), ("store D + fence + flag (L store follows)", (4, 1), (5, 1))]
    nbk = (len(Y) + 127) // 128
    print("\nchain step breakdown (us), median / mean over columns:")
    tot = 0.0
    for label, (ma, da), (mb, db) in names_m:
        v = [mk[(c + db, mb)] - mk[(c + da, ma)] for c in range(nbk - 1) if (c + da, ma) in mk and (c + db, mb) in mk]
        if v:
            print("  %-42s %7.2f %7.2f" % (label, float(np.median(v)), float(np.mean(v))))
            if not label.startswith("  "):
                tot += float(np.median(v))
    print("  %-42s %7.2f" % ("sum of medians", tot))

# ---- raw chain marks of one block column (DAG_TL_DUMP=c): every mark of columns c and c + 1, relative to the first
if os.environ.get("DAG_TL_DUMP") and len(marks):
    c0 = int(os.environ["DAG_TL_DUMP"])
    rows = [((tm - t0) * 1e-3, int((cd >> 20) & 0xfff), int((cd >> 8) & 0xff)) for tm, _, _, cd in marks
            if int((cd >> 20) & 0xfff) in (c0, c0 + 1) and int((cd >> 8) & 0xff) < 30 or
            (int((cd >> 20) & 0xfff) in (c0, c0 + 1) and 40 <= int((cd >> 8) & 0xff) < 64)]
    rows.sort()
    print("\nmarks of columns %d, %d (us after the first; column, mark):" % (c0, c0 + 1))
    for tm, c, m_ in rows:
        print("  %8.2f  c=%d  mark %d" % (tm - rows[0][0], c, m_))
