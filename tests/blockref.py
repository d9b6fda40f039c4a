"""numpy model of the DEVICE pipeline on a packed block (test helper, CPU).

It follows the engine's design, not the reference's: X and Y graphs are relabelled
*jointly*, features are (graph, column) counts, K[y, x] contracts all columns (a
column absent from X contributes 0), self similarities use all columns.  The CPU
tests compare it with the reference goldens to pin (a) the host packing logic and
(b) the joint-relabelling equivalence the C-ABI documents."""
from collections import Counter

import numpy as np
from scipy.sparse import csr_matrix


def _features(n_graphs, vgraph, cols, D):
    cnt = Counter(zip(vgraph.tolist(), cols.tolist()))
    r = np.fromiter((k[0] for k in cnt), dtype=np.int64, count=len(cnt))
    c = np.fromiter((k[1] for k in cnt), dtype=np.int64, count=len(cnt))
    d = np.fromiter(cnt.values(), dtype=np.float64, count=len(cnt))
    return csr_matrix((d, (r, c)), shape=(n_graphs, D))


def _finish(Phi, n_graphs, n_fit, normalize, nan_to_num, rows=None, n_rows=None):
    if rows is not None:  # gk_set_row_map: packed graph g feeds row rows[g] (disjoint columns per source)
        P = csr_matrix((np.ones(len(rows)), (np.asarray(rows), np.arange(len(rows)))), shape=(n_rows, Phi.shape[0]))
        Phi = csr_matrix(P.dot(Phi))
        n_graphs = n_rows
    X = Phi[:n_fit]
    diag = np.asarray(Phi.multiply(Phi).sum(axis=1)).ravel()
    if n_fit == n_graphs:
        K = X.dot(X.T).toarray()
        den = np.sqrt(np.outer(diag, diag))
    else:
        K = Phi[n_fit:].dot(X.T).toarray()
        den = np.sqrt(np.outer(diag[n_fit:], diag[:n_fit]))
    if normalize:
        with np.errstate(all="ignore"):
            K = K / den
            if nan_to_num:
                K = np.nan_to_num(K)
    return K, diag[:n_fit], diag[n_fit:]


def wl_levels(block, ids, n_iter):
    """Per-level dense labels (first-occurrence numbering) of every vertex."""
    V = block.n_vertices
    rp, ci = block.row_ptr, block.col_idx
    lab = np.asarray(ids, dtype=np.int64)
    levels = [lab.copy()]
    for _ in range(n_iter):
        table = {}
        new = np.empty(V, dtype=np.int64)
        for v in range(V):
            sig = (int(lab[v]), tuple(sorted(lab[ci[rp[v]:rp[v + 1]]].tolist())))
            new[v] = table.setdefault(sig, len(table))
        lab = new
        levels.append(lab.copy())
    return levels


def wl_gram_block(block, ids, n_iter, n_fit=None, normalize=False, rows=None, n_rows=None):
    N = block.n_graphs
    n_fit = (N if rows is None else n_rows) if n_fit is None else n_fit
    vgraph = np.repeat(np.arange(N), np.diff(block.graph_ptr))
    levels = wl_levels(block, ids, n_iter)
    cols, base = [], 0
    for lab in levels:
        cols.append(lab + base)
        base += int(lab.max()) + 1 if len(lab) else 0
    Phi = _features(N, np.tile(vgraph, len(levels)), np.concatenate(cols), base)
    return _finish(Phi, N, n_fit, normalize, True, rows, n_rows)


def wl_oa_gram_block(block, ids, n_iter, n_fit=None, normalize=False):
    """gk_wl_oa_features model: unary expansion of the WL count block -- the k-th vertex of a graph
    carrying column c lands in column (c, k) -- then the ordinary Gram (dot products of 0/1 rows =
    histogram intersections)."""
    N = block.n_graphs
    n_fit = N if n_fit is None else n_fit
    vgraph = np.repeat(np.arange(N), np.diff(block.graph_ptr))
    seen = Counter()
    enum = {}
    rows, cols = [], []
    base = 0
    for lab in wl_levels(block, ids, n_iter):
        for g, c in zip(vgraph.tolist(), (lab + base).tolist()):
            seen[(g, c)] += 1
            rows.append(g)
            cols.append(enum.setdefault((c, seen[(g, c)]), len(enum)))
        base += int(lab.max()) + 1 if len(lab) else 0
    Phi = _features(N, np.asarray(rows, dtype=np.int64), np.asarray(cols, dtype=np.int64), max(len(enum), 1))
    return _finish(Phi, N, n_fit, normalize, True)


def apsp_block(block, g):
    v0, v1 = int(block.graph_ptr[g]), int(block.graph_ptr[g + 1])
    n = v1 - v0
    D = np.full((n, n), np.inf)
    for u in range(n):
        b, e = block.row_ptr[v0 + u], block.row_ptr[v0 + u + 1]
        for k in range(b, e):
            w = 1.0 if block.weights is None else float(block.weights[k])
            D[u, block.col_idx[k] - v0] = w
    np.fill_diagonal(D, 0)
    for k in range(n):
        D = np.minimum(D, D[:, k:k + 1] + D[k:k + 1, :])
    return D


def sp_gram_block(block, ids, n_fit=None, with_labels=True, normalize=False, rows=None, n_rows=None, wl_iter=None,
                  nan_to_num=False):
    """ShortestPath features; with `wl_iter` the gk_wl_sp_features model: the labelled path
    histogram of every WL level with level-unique label ids in one feature block."""
    N = block.n_graphs
    n_fit = (N if rows is None else n_rows) if n_fit is None else n_fit
    if wl_iter is None:
        label_sets = [np.asarray(ids, dtype=np.int64)] if with_labels else [None]
    else:
        label_sets, base = [], 0
        for lab in wl_levels(block, ids, wl_iter):
            label_sets.append(lab + base)
            base += int(lab.max()) + 1 if len(lab) else 0
    enum = {}
    frows, cols = [], []
    for g in range(N):
        D = apsp_block(block, g)
        v0 = int(block.graph_ptr[g])
        n = D.shape[0]
        for lab in label_sets:
            for u in range(n):
                for v in range(n):
                    if u == v or not np.isfinite(D[u, v]):
                        continue
                    key = (int(lab[v0 + u]), int(lab[v0 + v]), D[u, v]) if lab is not None else D[u, v]
                    frows.append(g)
                    cols.append(enum.setdefault(key, len(enum)))
    Phi = _features(N, np.asarray(frows, dtype=np.int64), np.asarray(cols, dtype=np.int64), max(len(enum), 1))
    return _finish(Phi, N, n_fit, normalize, nan_to_num or wl_iter is not None, rows, n_rows)
