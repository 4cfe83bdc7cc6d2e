"""Generate the committed fixtures under tests/golden/ from the reference's bundled datasets.

Run in the BUILD container only (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
Outputs:
  * <dataset>.csv      - the reference data frames data/*.rda (R/data.R:1-115) as plain text
  * oracle_cases.json  - inputs/outputs of the CPU oracle (oracle/gpedm_oracle.py) for the README
                         cases K1-K5 (README.md:79-101, 254-312, 444-456, 515-571) at full double
                         precision, so GPU parity tests do not have to re-run the slow brute-force
                         parts of the oracle.
The README's printed digits themselves are hard-coded in tests/test_oracle_readme.py.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import gpedm_oracle as O  # noqa: E402
from oracle.rda import read_rda  # noqa: E402

REF = "/root/reference/data"


def write_csv(name, df):
    cols = list(df.keys())
    with open(os.path.join(HERE, name + ".csv"), "w") as f:
        f.write(",".join(cols) + "\n")
        n = len(df[cols[0]])
        for i in range(n):
            f.write(",".join(repr(float(df[c][i])) if df[c].dtype != object else str(df[c][i]) for c in cols) + "\n")


def tolist(v):
    if isinstance(v, dict):
        return {str(k): tolist(x) for k, x in v.items()}
    if isinstance(v, np.ndarray):
        return [None if (isinstance(x, float) and np.isnan(x)) else x for x in v.astype(object).tolist()] \
            if v.dtype != object else [str(x) for x in v]
    if isinstance(v, (np.floating, np.integer)):
        return v.item()
    return v


def pack(m, keys=("pars", "parst", "grad", "iter", "insampfitstats", "insampfitstatspop", "insampresults",
                  "outsampresults", "outsampfitstats", "outsampfitstatspop", "GPgrad")):
    return {k: tolist(m[k]) for k in keys if k in m}


def main():
    data = {}
    for name in ["thetalog2pop", "HastPow3sp", "shrimp", "trawl", "RickerHarvest"]:
        df = read_rda(os.path.join(REF, name + ".rda"))[name]
        write_csv(name, df)
        data[name] = df
    tl = data["thetalog2pop"]
    cases = {}
    # K1
    m = O.fitGP(data=tl, y="Abundance", pop="Population", E=3, tau=1, scaling="local", predictmethod="loo")
    cases["K1"] = pack(m)
    cases["K1"]["lto"] = pack(O.predict(m, predictmethod="lto"))
    cases["K1"]["lto_r1"] = pack(O.predict(m, predictmethod="lto", exclradius=1))
    cases["K1"]["loo_r2"] = pack(O.predict(m, predictmethod="loo", exclradius=2))
    cases["K1"]["loo_grad"] = pack(O.predict(m, predictmethod="loo", returnGPgrad=True))
    cases["K1"]["lto_grad"] = pack(O.predict(m, predictmethod="lto", returnGPgrad=True))
    cases["K1"]["sequential"] = pack(O.predict(m, predictmethod="sequential"))
    # K5
    m = O.fitGP(y=tl["Abundance"], pop=tl["Population"], E=2, tau=1, scaling="local")
    cases["K5"] = pack(m)
    # K2 / K3
    pA = {k: v[tl["Population"] == "PopA"] for k, v in tl.items()}
    lags = O.makelags_subrt(pA["Abundance"], pA["Population"], 2, 1)
    pAd = dict(pA)
    pAd["Abundance_1"], pAd["Abundance_2"] = lags[:, 0], lags[:, 1]
    tr = {k: v[:40] for k, v in pAd.items()}
    te = {k: v[40:50] for k, v in pAd.items()}
    m3 = O.fitGP(data=tr, y="Abundance", x=["Abundance_1", "Abundance_2"], time="Time", newdata=te)
    cases["K2"] = pack(m3)
    cases["K2"]["sequential"] = pack(O.predict(m3, predictmethod="sequential"))
    m3u = O.predict_seq(m3, tr, te)
    cases["K3"] = pack(m3u)
    # K4
    lags1 = O.makelags_subrt(tl["Abundance"], tl["Population"], 3, 1)
    fore1 = O.makelags_subrt(tl["Abundance"], tl["Population"], 3, 1, forecast=True)
    d1 = dict(tl)
    for i in range(3):
        d1["Abundance_%d" % (i + 1)] = lags1[:, i]
    tl2 = O.fitGP(data=d1, y="Abundance", x=["Abundance_1", "Abundance_2", "Abundance_3"], pop="Population",
                  time="Time", scaling="local")
    nd = {"Time": np.array([51., 51.]), "Population": np.array(["PopA", "PopB"], dtype=object)}
    for i in range(3):
        nd["Abundance_%d" % (i + 1)] = fore1[:, i]
    cases["K4"] = pack(tl2)
    cases["K4"]["forecast"] = pack(O.predict(tl2, newdata=nd, returnGPgrad=True))
    cases["K4"]["fore1"] = fore1.tolist()
    with open(os.path.join(HERE, "oracle_cases.json"), "w") as f:
        json.dump(cases, f, indent=0)
    print("wrote", sorted(cases))


if __name__ == "__main__":
    main()
