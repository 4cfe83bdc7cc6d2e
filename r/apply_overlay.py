#!/usr/bin/env python3
"""Apply r/predict_GP.ed to a copy of the reference's R/GPfunctions.R.

    python r/apply_overlay.py /path/to/GPEDM/R/GPfunctions.R /path/to/patched/GPfunctions.R

(`ed` / `patch -e` do the same where they are installed.)  Before editing, every block the script
removes is fingerprinted: the first and last removed line must contain the expected token, so a
different upstream revision is refused instead of being silently mis-patched."""
import re
import sys

# original line -> token that line must contain
ANCHORS = {460: "sqrt(diag(output$cov))", 461: "output$pars[\"ve\"]", 884: "iKVs=object$covm$iKVs",
           1012: "#get covariance matrix", 1021: "}", 1025: "GPgrad=Xnew*NA", 1030: "}",
           1057: "Cd=object$covm$Cd", 1058: "Sigma=object$covm$Sigma", 1067: "for(i in 1:nd)", 1086: "}",
           1144: "Cd=object$covm$Cd", 1145: "Sigma=object$covm$Sigma", 1153: "for(i in 1:nt)", 1176: "}",
           1203: "S=object$covm$Cd", 1215: "ymean1=matrix(0", 1216: "ysd1=(sigma2+ve)", 1223: "for(i in 1:(nt-1))",
           1245: "}"}


def parse(script):
    """[(first, last, 'c'|'d', [new lines])] in script order (bottom-up)."""
    cmds, lines, i = [], script.splitlines(), 0
    while i < len(lines):
        ln = lines[i]
        i += 1
        if not ln.strip() or ln.startswith("#"):
            continue
        m = re.fullmatch(r"(\d+)(?:,(\d+))?([cd])", ln.strip())
        if not m:
            raise ValueError("cannot parse ed command %r" % ln)
        a, b, op = int(m.group(1)), int(m.group(2) or m.group(1)), m.group(3)
        new = []
        if op == "c":
            while lines[i] != ".":
                new.append(lines[i])
                i += 1
            i += 1
        cmds.append((a, b, op, new))
    return cmds


def apply(src_text, script):
    src = src_text.splitlines()
    for ln, tok in ANCHORS.items():
        if ln > len(src) or tok not in src[ln - 1]:
            raise SystemExit("anchor mismatch at line %d: expected %r, found %r - not the surveyed revision of "
                             "R/GPfunctions.R" % (ln, tok, src[ln - 1] if ln <= len(src) else None))
    last = len(src) + 1
    for a, b, op, new in parse(script):
        if not (1 <= a <= b < last):
            raise SystemExit("edit script is not bottom-up at %d,%d" % (a, b))
        src[a - 1:b] = new
        last = a
    return "\n".join(src) + "\n"


if __name__ == "__main__":
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    out = apply(open(sys.argv[1]).read(), open(os.path.join(here, "predict_GP.ed")).read())
    open(sys.argv[2], "w").write(out)
    print("patched file written to", sys.argv[2])
