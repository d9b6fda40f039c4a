"""Golden fixtures for WeisfeilerLehmanOptimalAssignment, produced by the
REAL reference.  Inputs are regenerated from seeds, only the reference's matrices are stored
(tests/golden/wloa.npz).

    GRAKEL_REF=/tmp/grakel_ref python tests/golden/make_golden_oa.py
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.environ.get("GRAKEL_REF", "/tmp/ref"))

import gio  # noqa: E402

from grakel import WeisfeilerLehmanOptimalAssignment as RefOA  # noqa: E402  (the reference)

from oracle.gk_oracle import gen, oa_sets  # noqa: E402

warnings.simplefilter("ignore")
out = {}
fit, new = oa_sets()
for tag, nrm in (("", False), ("_n", True)):
    for h in (1, 3):
        k = RefOA(n_iter=h, normalize=nrm)
        out[f"oa_h{h}_K{tag}"] = k.fit_transform(fit)
        out[f"oa_h{h}_Kt{tag}"] = k.transform(new)
        xd, yd = k.diagonal()
        out[f"oa_h{h}_xd{tag}"] = np.asarray(xd, dtype=float)
        out[f"oa_h{h}_yd{tag}"] = np.asarray(yd, dtype=float)
k = RefOA(n_iter=2, sparse=True)
out["oa_sparse_K"] = k.fit_transform(fit)
M = gio.dec_dataset(gio.load(os.path.join(HERE, "mutag_graphs.json.gz")))[:60]
out["oa_mutag_h4"] = RefOA(n_iter=4).fit_transform(M)
out["oa_cfg1_h3"] = RefOA(n_iter=3).fit_transform(gen(188, 18, 0)[:80])
np.savez_compressed(os.path.join(HERE, "wloa.npz"), **out)
for k, v in out.items():
    print(k, v.shape, float(np.nansum(v)))
