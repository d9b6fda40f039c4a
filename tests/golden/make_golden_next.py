"""Golden fixtures for the SURVEY 8(f) "next" rows, produced by the REAL reference:
EdgeHistogram, WeisfeilerLehman over EdgeHistogram / ShortestPath, CoreFramework over
WeisfeilerLehman / ShortestPath.  Inputs are regenerated from seeds (oracle generators), so
only the reference's matrices are stored (tests/golden/next_rows.npz).

    GRAKEL_REF=/tmp/grakel_ref python tests/golden/make_golden_next.py
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

from grakel import CoreFramework, EdgeHistogram, ShortestPath, WeisfeilerLehman  # noqa: E402  (the reference)

from oracle.gk_oracle import gen, gen_edge_labelled  # noqa: E402

warnings.simplefilter("ignore")
out = {}
X = gen_edge_labelled(36, 10, 3)
fit, new = X[:26], X[26:]
for tag, nrm in (("", False), ("_n", True)):
    e = EdgeHistogram(normalize=nrm)
    out["eh_K" + tag] = e.fit_transform(fit)
    out["eh_Kt" + tag] = e.transform(new)
    for name, cls in (("wleh", EdgeHistogram), ("wlsp", ShortestPath)):
        w = WeisfeilerLehman(n_iter=2, base_graph_kernel=cls, normalize=nrm)
        out[f"{name}_K{tag}"] = w.fit_transform(fit)
        out[f"{name}_Kt{tag}"] = w.transform(new)
# deeper WL-SP on the bundled MUTAG prefix (dictionary inputs with string-free labels)
M = gio.dec_dataset(gio.load(os.path.join(HERE, "mutag_graphs.json.gz")))[:48]
out["wlsp_mutag_h3"] = WeisfeilerLehman(n_iter=3, base_graph_kernel=ShortestPath).fit_transform(M)
# CoreFramework (adjacency inputs; denser graphs so that cores above 1 exist)
C = gen(30, 12, 5, as_adj=True)
rs = np.random.RandomState(7)
for el in C:  # add random extra edges: average degree ~6
    A = el[0]
    n = A.shape[0]
    extra = np.triu(rs.rand(n, n) < 0.25, 1)
    A[:] = ((A + extra + extra.T) > 0).astype(float)
cfit, cnew = C[:22], C[22:]
for tag, nrm in (("", False), ("_n", True)):
    for name, base in (("corewl", (WeisfeilerLehman, {"n_iter": 2})), ("coresp", ShortestPath)):
        c = CoreFramework(base_graph_kernel=base, normalize=nrm)
        out[f"{name}_K{tag}"] = c.fit_transform(cfit)
        out[f"{name}_Kt{tag}"] = c.transform(cnew)
c = CoreFramework(base_graph_kernel=(WeisfeilerLehman, {"n_iter": 2}), min_core=1)
out["corewl_min1_K"] = c.fit_transform(cfit)
np.savez_compressed(os.path.join(HERE, "next_rows.npz"), **out)
for k, v in out.items():
    print(k, v.shape, float(np.nansum(v)))
