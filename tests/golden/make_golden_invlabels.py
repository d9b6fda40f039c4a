"""Golden fixture for the per-level label dictionaries of a fitted WeisfeilerLehman (`_inv_labels`,
weisfeiler_lehman.py:208-257), produced by the REAL reference (baseline/_ref, or GRAKEL_REF).  Inputs are
regenerated from the seed (oracle.gk_oracle.gen); only the reference's dictionaries are stored.

    python tests/golden/make_golden_invlabels.py
"""
import gzip
import json
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..", "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.environ.get("GRAKEL_REF", os.path.join(ROOT, "baseline", "_ref")))

from grakel import WeisfeilerLehman  # noqa: E402  (the reference)

from oracle.gk_oracle import gen  # noqa: E402

CASES = {"er60": dict(N=60, nbar=14, seed=77, nl=3, n_iter=3),
         "er25_deep": dict(N=25, nbar=9, seed=5, nl=2, n_iter=5)}

if __name__ == "__main__":
    warnings.simplefilter("ignore")
    out = {}
    for name, c in CASES.items():
        X = gen(c["N"], c["nbar"], c["seed"], nl=c["nl"])
        wl = WeisfeilerLehman(n_iter=c["n_iter"])
        wl.fit(X)
        out[name] = {"params": c, "levels": {str(i): sorted([[k, int(v)] for k, v in d.items()], key=lambda kv: kv[1])
                                             for i, d in wl._inv_labels.items()}}
        print(name, {i: len(d) for i, d in wl._inv_labels.items()})
    with gzip.open(os.path.join(HERE, "inv_labels.json.gz"), "wt") as f:
        json.dump(out, f)
