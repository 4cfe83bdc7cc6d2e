"""Hand-off variants of the Cholesky chain inside the persistent factorisation kernel (csrc/dag_factor.cuh; replaces
LAPACK dpotrf behind Matrix::chol, reference R/GPfunctions.R:826-829).

  strips (GPEDM_DAG_NO_STRIPS)   L(c+1, c) is handed to DIAG(c+1) in 32-column strips - timing only, same arithmetic
  solve  (GPEDM_DAG_TRSM = t)    L(c+1 .. c+t, c) solved against the block columns of L_cc (forward substitution)
                                 instead of multiplied by D_c = L_cc^-1 - same result up to rounding

The knobs are read when the library is loaded, so every variant runs in a child process."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CODE = """
import json, sys
sys.path.insert(0, %r)
import numpy as np
import gpedm_b200 as G
from gpedm_b200 import synth
from gpedm_b200.batch import default_initparst
out = []
for n, E in [(260, 3), (1024, 10), (2059, 10)]:
    Y, X = synth.embed(synth.ricker(n, 1004), E, 1)
    m = G.GPModel(Y, X)
    p0 = default_initparst(E, True)
    r1 = m.likegrad(p0, 1.0, None, None, True)
    r2 = m.likegrad(p0, 1.0, None, None, True)
    out.append(dict(N=len(Y), nll=float(r1["nllpost"]).hex(), grad=[float(v).hex() for v in r1["grad"]],
                    same=bool(float(r1["nllpost"]) == float(r2["nllpost"]) and np.array_equal(r1["grad"], r2["grad"]))))
    m.close()
print(json.dumps(out))
""" % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra):
    r = subprocess.run([sys.executable, "-c", CODE], capture_output=True, text=True, env=dict(os.environ, **extra), timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_chain_handoff_variants_agree(gpu_lib):
    default = _run({})
    product = _run({"GPEDM_DAG_TRSM": "0"})          # strips, every tile multiplied by D_c
    whole = _run({"GPEDM_DAG_NO_STRIPS": "1"})       # round-2 v5 behaviour: whole tile + flag, product with D_c
    solve1 = _run({"GPEDM_DAG_TRSM": "1"})
    assert [o["N"] for o in default] == [257, 1014, 2049]
    for o in default + product + whole + solve1:
        assert o["same"], "repeated evaluations must be bit-identical (fixed summation order, no timing dependence)"
    assert product == whole, "the strips change when a tile is read, never what is computed"
    for d, s1, p in zip(default, solve1, product):
        a, c, b = float.fromhex(d["nll"]), float.fromhex(s1["nll"]), float.fromhex(p["nll"])
        assert abs(a - b) <= 1e-11 * abs(b) and abs(c - b) <= 1e-11 * abs(b)   # north-star tolerance on nll: 1e-10
        ga, gb = np.array([float.fromhex(v) for v in d["grad"]]), np.array([float.fromhex(v) for v in p["grad"]])
        assert np.max(np.abs(ga - gb)) <= 1e-9 * max(1.0, np.max(np.abs(gb)))   # gradient tolerance: 1e-8
