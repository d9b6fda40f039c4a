"""Golden fixtures for ShortestPathAttr with a USER metric (shortest_path.py:130-164 through the generic pairwise driver
kernel.py:236-296), produced by the REAL reference (baseline/_ref, or GRAKEL_REF): unit weights (adjacency input,
Floyd-Warshall), real-valued weights (edge dictionaries, Dijkstra), fit_transform + transform, normalised and not.
Inputs are regenerated from the seed (tests share `gen_attr` / `rbf` below); only the reference's matrices are stored.

    python tests/golden/make_golden_spattr_metric.py
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..", "..")


def rbf(a, b):
    """A non-bilinear metric: Gaussian of the attribute difference."""
    d = np.asarray(a, dtype=float) - np.asarray(b, dtype=float)
    return float(np.exp(-0.5 * np.dot(d, d)))


def gen_attr(n_graphs, nbar, seed, dim=3, real_weights=False):
    """Small ER graphs with attribute vectors; adjacency matrices (unit weights) or {(u, v): w} dictionaries."""
    rs = np.random.RandomState(seed)
    out = []
    for _ in range(n_graphs):
        n = int(rs.randint(max(2, nbar // 2), nbar + nbar // 2 + 1))
        iu = np.triu_indices(n, 1)
        m = rs.rand(len(iu[0])) < 2.5 / max(n - 1, 1)
        L = {i: rs.rand(dim).round(3).tolist() for i in range(n)}
        if real_weights:
            g = {}
            for a, b in zip(iu[0][m].tolist(), iu[1][m].tolist()):
                w = float(rs.choice([0.1, 0.2, 0.3, 0.7]))
                g[(a, b)] = w
                g[(b, a)] = w
            if not g:
                g[(0, 1)] = g[(1, 0)] = 0.5
            out.append([g, L])
        else:
            A = np.zeros((n, n), dtype=int)
            A[iu[0][m], iu[1][m]] = 1
            A = A + A.T
            out.append([A.tolist(), L])
    return out


if __name__ == "__main__":
    sys.path.insert(0, os.environ.get("GRAKEL_REF", os.path.join(ROOT, "baseline", "_ref")))
    from grakel import ShortestPathAttr  # noqa: E402  (the reference)

    warnings.simplefilter("ignore")
    out = {}
    X = gen_attr(9, 7, 11)
    fit, new = X[:6], X[6:]
    e = ShortestPathAttr(metric=rbf)
    out["unit_K"] = e.fit_transform(fit)
    out["unit_Kt"] = e.transform(new)
    e = ShortestPathAttr(metric=rbf, normalize=True)
    out["unit_Kn"] = e.fit_transform(fit)
    out["unit_Ktn"] = e.transform(new)
    W = gen_attr(6, 6, 12, real_weights=True)
    e = ShortestPathAttr(metric=rbf)
    out["real_K"] = e.fit_transform(W[:4])
    out["real_Kt"] = e.transform(W[4:])
    np.savez_compressed(os.path.join(HERE, "spattr_metric.npz"), **out)
    for k, v in out.items():
        print(k, v.shape, float(np.nansum(v)))
