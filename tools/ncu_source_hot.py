"""Warp-stall samples of one kernel in an .ncu-rep, by opcode and by CUDA source line.

    python tools/ncu_source_hot.py gpurun_out/x.ncu-rep [top]

Reads `ncu -i REP --page source --csv --print-source cuda,sass` (needs a capture taken with
`--set full --import-source on` of a `-lineinfo` build) and prints (1) the share of samples per SASS opcode,
(2) the source lines ranked by samples with their dominant stall reasons, (3) the instructions executed per
opcode.  Used to find what the warps of the persistent factorisation kernel do when they are not issuing DMMAs."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur, hdr, lines, instrs, curline = None, None, [], [], ('', '')
for r in rows:
    if r and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr):
        if r[2] == "-":  # a source line; the rows of its SASS instructions follow
            curline = (r[0], r[1].strip())
            lines.append((cur, r))
        else:
            instrs.append((cur, r, curline))
ix = {}
for i, h in enumerate(hdr):
    ix.setdefault(h, i)
S, EXE = ix["# Samples"], ix["Instructions Executed"]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


def opcode(src):
    p = src.split()
    return p[1] if p and p[0].startswith("@") else (p[0] if p else "?")


tot = sum(num(r[S]) for _, r, _l in instrs)
by_op, ex_op = {}, {}
for _, r, _l in instrs:
    op = opcode(r[3])
    by_op[op] = by_op.get(op, 0) + num(r[S])
    ex_op[op] = ex_op.get(op, 0) + num(r[EXE])
print("samples %d over %d SASS instructions" % (tot, len(instrs)))
print("-- samples by opcode")
for k, v in sorted(by_op.items(), key=lambda x: -x[1])[:14]:
    print("  %-32s %6.2f %%" % (k, 100.0 * v / tot))
agg = {s: sum(num(r[ix[s]]) for _, r, _l in instrs) for s in stalls}
print("-- stall reasons: " + ", ".join("%s %.1f" % (k[6:], 100.0 * v / tot) for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v > 0.003 * tot))
# per source line: sum over its SASS instructions (the summary rows of multi-line asm statements are unreliable)
per_line = {}
for f, r, cl in instrs:
    key = (f, cl[0], cl[1])
    d = per_line.setdefault(key, [0] + [0] * len(stalls))
    d[0] += num(r[S])
    for q, s in enumerate(stalls):
        d[1 + q] += num(r[ix[s]])
print("-- source lines by samples (%% of all samples; dominant stalls)")
for (f, ln, src), d in sorted(per_line.items(), key=lambda x: -x[1][0])[:top]:
    dom = sorted(((d[1 + q], stalls[q][6:]) for q in range(len(stalls))), reverse=True)[:2]
    print("  %-18s %5s %6.2f  %-22s %s" % (f[:18], ln, 100.0 * d[0] / tot, " ".join("%s:%d" % (n, 100 * v // max(1, d[0])) for v, n in dom), src[:84]))
print("-- instructions executed by opcode")
te = sum(ex_op.values())
for k, v in sorted(ex_op.items(), key=lambda x: -x[1])[:12]:
    print("  %-32s %14d %6.2f %%" % (k, v, 100.0 * v / te))
