"""SM-occupancy timeline of ONE evaluation from the per-CTA log of a -DGPEDM_TIMELINE build.

  nvcc ... -DGPEDM_TIMELINE -o build/libgpedm_timeline.so gpedm_b200/csrc/gpedm_core.cu -ldl
  python tools/timeline.py [series_len] [bin_us]

Every CTA of the tile kernels and of the diagonal-block kernel logs (globaltimer start, end, SM id, op | tag << 8).
Printed: per time bin, the fraction of the 148 SMs busy per category, and the totals (SM-ms per category, idle SM-ms)."""
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
L.gpedm_debug_timeline(None, 0, C.byref(cnt))          # reset
m.likegrad(p0, 1.0, None, None, True)
MAXR = 1 << 18
buf = np.zeros(4 * MAXR, dtype=np.uint64)
L.gpedm_debug_timeline(buf.ctypes.data_as(C.POINTER(C.c_ulonglong)), MAXR, C.byref(cnt))
nrec = cnt.value
rec = buf[:4 * nrec].reshape(nrec, 4).astype(np.int64)
t0 = rec[:, 0].min()
st, en, sm, code = (rec[:, 0] - t0) * 1e-3, (rec[:, 1] - t0) * 1e-3, rec[:, 2], rec[:, 3]   # us
typ, tag = code & 0xff, (code >> 8) & 0xff
flop = 2.0 * ((code >> 40) & 0xfff) * ((code >> 52) & 0xfff) * ((code >> 16) & 0xffffff) * 16   # 2 TM TN (nslab BK)
names = {"diag (chain)": (typ == 100), "panel (chain)": (typ == 0), "column updates (chain)": (typ == 1) & (tag == 1),
         "bulk trailing update": (typ == 1) & (tag == 0), "trtri early (fill stream)": ((typ == 2) | (typ == 3)) & (tag == 2),
         "trtri late (main stream)": ((typ == 2) | (typ == 3)) & (tag != 2), "lauum": (typ == 4)}
nsm = int(sm.max()) + 1
T = en.max()
phases = m.phase_ms()
print("N=%d  records=%d  SMs seen=%d  span %.3f ms (tile + diag kernels only; assemble/trace/vector kernels not logged)" % (len(Y), nrec, nsm, T * 1e-3))
print("phase events (ms):", {k: round(v, 3) for k, v in phases.items()})
nb = int(np.ceil(T / bin_us))
edges = np.arange(nb + 1) * bin_us
tot = {}
rows = {}
for k, mask in names.items():
    s_, e_ = st[mask], en[mask]
    tot[k] = float(np.sum(e_ - s_))
    occ = np.zeros(nb)
    for b in range(nb):
        occ[b] = np.sum(np.clip(np.minimum(e_, edges[b + 1]) - np.maximum(s_, edges[b]), 0, None))
    rows[k] = occ / (nsm * bin_us)
print("\nbusy fraction of the %d SMs per %.0f us bin (CTA residency; 128x64 half-tile launches hold 2 CTAs per SM and count twice)" % (nsm, bin_us))
hdr = ["t_ms"] + [k.split(" (")[0][:12] for k in names] + ["total"]
print(" ".join("%12s" % h for h in hdr))
for b in range(nb):
    vals = [rows[k][b] for k in names]
    print(" ".join(["%12.2f" % (edges[b] * 1e-3)] + ["%12.2f" % v for v in vals] + ["%12.2f" % sum(vals)]))
print("\nSM-ms per category (sum of CTA lifetimes / 1000):")
for k in names:
    gf = float(np.sum(flop[names[k]])) * 1e-9
    print("  %-28s %9.2f SM-ms  = %5.2f ms of the whole GPU   %7.1f GFLOP executed -> %5.1f TFLOP/s while resident"
          % (k, tot[k] * 1e-3, tot[k] * 1e-3 / nsm, gf, gf / max(tot[k] * 1e-6 / nsm, 1e-12) * 1e-3))
print("  (executed flops count whole 128-wide tiles; the N^3 model of the evaluation is %.1f GFLOP)" % (float(len(Y)) ** 3 * 1e-9))
busy = sum(tot.values()) * 1e-3
print("  %-28s %9.2f SM-ms  = %5.2f ms of the whole GPU" % ("idle (span x SMs - busy)", nsm * T * 1e-3 - busy, T * 1e-3 - busy / nsm))
# chain kernels: latency view
for k in ("diag (chain)", "panel (chain)", "column updates (chain)"):
    mask = names[k]
    if mask.any():
        print("  %-28s %d CTAs, mean CTA lifetime %.1f us" % (k, int(mask.sum()), float(np.mean((en - st)[mask]))))

# ---- chain view: union of the intervals covered by chain CTAs (diag, panel, aux-stream updates) vs the gaps
chain = (typ == 100) | (typ == 0) | ((typ == 1) & (tag == 1))
iv = sorted(zip(st[chain], en[chain]))
cov, gaps, cur_s, cur_e = 0.0, [], iv[0][0], iv[0][1]
for a, b in iv[1:]:
    if a > cur_e:
        cov += cur_e - cur_s
        gaps.append(a - cur_e)
        cur_s, cur_e = a, b
    else:
        cur_e = max(cur_e, b)
cov += cur_e - cur_s
gaps = np.array(gaps)
print("\nchain view: first chain CTA starts at %.1f us, last ends at %.1f us; covered by >= 1 chain CTA %.1f us;"
      " %d gaps with no chain CTA resident, total %.1f us (median %.2f us, max %.1f us)"
      % (iv[0][0], cur_e, cov, len(gaps), gaps.sum(), float(np.median(gaps)) if len(gaps) else 0.0, gaps.max() if len(gaps) else 0.0))
big = []
cur_e = iv[0][1]
for a, b in iv[1:]:
    if a > cur_e and a - cur_e >= 15.0:
        big.append((round(cur_e, 1), round(a - cur_e, 1)))
    cur_e = max(cur_e, b)
print("gaps >= 15 us (start us, length us):", big)
