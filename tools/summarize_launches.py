"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per-kernel totals and shares.
usage: python tools/summarize_launches.py launches.csv [first_launch last_launch]   (window = ids to keep)"""
import csv, sys, re, collections
rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        rows.append((int(r["ID"]), r["Kernel Name"], float(r["Metric Value"].replace(",", "")), r["Metric Unit"]))
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (0, 1 << 60)
rows = [r for r in rows if lo <= r[0] < hi]
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}
tot = collections.OrderedDict()
for _, name, v, u in rows:
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"gpedm::", "", name)
    t = tot.setdefault(name, [0.0, 0])
    t[0] += v * scale.get(u, 1e-6)
    t[1] += 1
total = sum(v[0] for v in tot.values())
print("launch ids [%d, %d): %d launches, %.3f ms serialised" % (rows[0][0], rows[-1][0] + 1, len(rows), total))
for name, (ms, n) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print("%9.3f ms %5.1f%%  x%4d  %s" % (ms, 100 * ms / total, n, name))
