"""WeisfeilerLehmanOptimalAssignment.  Goldens come from the REAL reference
(tests/golden/make_golden_oa.py -> tests/golden/wloa.npz).

CPU (-m "not gpu"): the oracle restatement against the goldens; the host logic of the estimator (the
"keys of the edge dictionary" vertex set, joint relabelling, unary expansion) through the numpy model
of the device pipeline (tests/blockref.py).  GPU (-m gpu): the CUDA path through the C-ABI."""
import os
import warnings

import numpy as np
import pytest

import blockref
import gio
from grakel_b200 import GraphKernel, WeisfeilerLehmanOptimalAssignment
from grakel_b200.packing import Block, pack
from oracle.gk_oracle import WLOAOracle, gen, oa_sets

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")
REF = np.load(os.path.join(G, "wloa.npz"))


def _eq(a, b, exact=True):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    assert a.shape == b.shape
    if exact:
        assert np.array_equal(a, b, equal_nan=True), f"max abs diff {np.nanmax(np.abs(a - b))}"
    else:
        np.testing.assert_allclose(a, b, rtol=1e-12, atol=0, equal_nan=True)


def _mutag():
    return gio.dec_dataset(gio.load(os.path.join(G, "mutag_graphs.json.gz")))[:60]


# ------------------------------------------------------------------ oracle vs the real reference
@pytest.mark.parametrize("tag,nrm", [("", False), ("_n", True)])
@pytest.mark.parametrize("h", [1, 3])
def test_oracle_matches_reference(tag, nrm, h):
    fit, new = oa_sets()
    o = WLOAOracle(n_iter=h, normalize=nrm)
    _eq(o.fit_transform(fit), REF[f"oa_h{h}_K{tag}"], exact=not nrm)
    _eq(o.transform(new), REF[f"oa_h{h}_Kt{tag}"], exact=not nrm)
    _eq(o.xdiag, REF[f"oa_h{h}_xd{tag}"])
    _eq(o.ydiag, REF[f"oa_h{h}_yd{tag}"])


def test_oracle_matches_reference_deeper_sets():
    fit, _ = oa_sets()
    _eq(WLOAOracle(n_iter=2).fit_transform(fit), REF["oa_sparse_K"])  # `sparse=True` only changes the storage
    _eq(WLOAOracle(n_iter=4).fit_transform(_mutag()), REF["oa_mutag_h4"])
    _eq(WLOAOracle(n_iter=3).fit_transform(gen(188, 18, 0)[:80]), REF["oa_cfg1_h3"])


# ------------------------------------------------------------------ host logic through the numpy device model
def _fitted(est, fit):
    est._method_calling = 1
    est.initialize()
    return est.parse_input(fit)


@pytest.mark.parametrize("tag,nrm", [("", False), ("_n", True)])
@pytest.mark.parametrize("h", [1, 3])
def test_host_logic_and_unary_expansion(tag, nrm, h):
    fit, new = oa_sets()
    est = WeisfeilerLehmanOptimalAssignment(n_iter=h, normalize=nrm)
    fx = _fitted(est, fit)
    est.X = fx
    K, xd, _ = blockref.wl_oa_gram_block(fx.block, fx.ids, h, normalize=nrm)
    _eq(K, REF[f"oa_h{h}_K{tag}"], exact=not nrm)
    _eq(xd, REF[f"oa_h{h}_xd{tag}"])
    est._method_calling = 3
    fy = est.parse_input(new)
    blk = Block.concat(fx.block, fy.block)
    Kt, _, yd = blockref.wl_oa_gram_block(blk, np.concatenate([fx.ids, fy.ids]), h, n_fit=fx.block.n_graphs, normalize=nrm)
    _eq(Kt, REF[f"oa_h{h}_Kt{tag}"], exact=not nrm)
    _eq(yd, REF[f"oa_h{h}_yd{tag}"])


def test_vertex_set_is_the_edge_dictionary():
    # a labelled vertex without any edge never reaches the histogram (dictionary input) ...
    b = pack([[{(0, 1): 1, (1, 0): 1}, {0: "a", 1: "b", 2: "c"}]], "wloa", len_ok=lambda n: n >= 2)
    assert b.n_vertices == 2 and sorted(b.labels) == ["a", "b"]
    # ... a sink does, and so does every row of an adjacency matrix
    b = pack([[{(0, 1): 1}, {0: "a", 1: "b"}]], "wloa", len_ok=lambda n: n >= 2)
    assert b.n_vertices == 2 and len(b.col_idx) == 1
    b = pack([[np.array([[0, 1, 0], [1, 0, 0], [0, 0, 0.0]]), {0: "a", 1: "b", 2: "c"}]], "wloa", len_ok=lambda n: n >= 2)
    assert b.n_vertices == 3
    with pytest.raises(KeyError):  # a vertex of an edge without a label: L[j][v] in the reference (:181)
        pack([[{(0, 1): 1}, {0: "a"}]], "wloa", len_ok=lambda n: n >= 2)


def test_error_behaviour_and_dispatch():
    fit, _ = oa_sets()
    with pytest.raises(TypeError):
        WeisfeilerLehmanOptimalAssignment(n_iter=0).initialize()
    with pytest.raises(TypeError):
        WeisfeilerLehmanOptimalAssignment(n_iter=2.0).initialize()
    e = WeisfeilerLehmanOptimalAssignment(n_iter=2)
    with pytest.raises(ValueError):
        e.fit_transform(None)
    e._method_calling = 1
    e.initialize()
    with pytest.raises(TypeError):
        e.parse_input(5)
    with pytest.raises(TypeError):
        e.parse_input([[fit[0][0]]])  # one member only (:113-135)
    with pytest.raises(ValueError):
        e.parse_input([[]])  # nothing left after the empty-element warning (:139-140)
    e.X = e.parse_input(fit)
    e._method_calling = 3
    with pytest.raises(ValueError):
        e.parse_input([[fit[0][0]]])  # transform raises ValueError for malformed elements (:344-346)
    with pytest.raises(ValueError):
        e.transform(None)
    for name in ("WL-OA", "weisfeiler_lehman_optimal_assignment"):
        gk = GraphKernel(kernel={"name": name, "n_iter": 3}, normalize=True)
        gk.initialize()
        assert isinstance(gk.kernel_, WeisfeilerLehmanOptimalAssignment)
        assert gk.kernel_.n_iter == 3 and gk.kernel_.normalize is True


# ------------------------------------------------------------------ the CUDA path
@pytest.mark.gpu
@pytest.mark.parametrize("tag,nrm", [("", False), ("_n", True)])
@pytest.mark.parametrize("h", [1, 3])
def test_gpu_wloa_matches_reference(tag, nrm, h):
    fit, new = oa_sets()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        e = WeisfeilerLehmanOptimalAssignment(n_iter=h, normalize=nrm)
        _eq(e.fit_transform(fit), REF[f"oa_h{h}_K{tag}"], exact=not nrm)
        _eq(e.transform(new), REF[f"oa_h{h}_Kt{tag}"], exact=not nrm)
        xd, yd = e.diagonal()
        _eq(xd, REF[f"oa_h{h}_xd{tag}"])
        _eq(yd, REF[f"oa_h{h}_yd{tag}"])
        f = WeisfeilerLehmanOptimalAssignment(n_iter=h).fit(fit)  # fit alone, then the diagonal
        _eq(f.diagonal(), REF[f"oa_h{h}_xd"])


@pytest.mark.gpu
def test_gpu_wloa_deeper_sets_and_dispatch():
    fit, _ = oa_sets()
    _eq(WeisfeilerLehmanOptimalAssignment(n_iter=2, sparse=True).fit_transform(fit), REF["oa_sparse_K"])
    _eq(WeisfeilerLehmanOptimalAssignment(n_iter=4).fit_transform(_mutag()), REF["oa_mutag_h4"])
    _eq(GraphKernel(kernel={"name": "WL-OA", "n_iter": 3}).fit_transform(gen(188, 18, 0)[:80]), REF["oa_cfg1_h3"])


@pytest.mark.gpu
@pytest.mark.parametrize("force_T", [None, "1", "4096"])
def test_gpu_wloa_against_the_oracle_on_every_gram_path(force_T, monkeypatch):
    """300 graphs, h = 4, fit + transform: default head/tail split, everything on the tensor cores
    (T = 1) and everything through the pair-update tail (T = 4096)."""
    if force_T is not None:
        monkeypatch.setenv("GRAKEL_B200_FORCE_T", force_T)
    X = gen(300, 16, 21, nl=3)
    fit, new = X[:220], X[220:]
    o = WLOAOracle(n_iter=4)
    e = WeisfeilerLehmanOptimalAssignment(n_iter=4)
    _eq(e.fit_transform(fit), o.fit_transform(fit))
    _eq(e.transform(new), o.transform(new))
    _eq(e.diagonal()[1], o.ydiag)


@pytest.mark.gpu
def test_gpu_wloa_full_size_properties():
    """BASELINE config 2's graphs (10 000, h = 5): the matrix is far beyond the oracle, so check what
    does not depend on the size -- K[i, j] only depends on graphs i and j (WL colours are canonical),
    hence any sub-matrix equals the oracle run on that subset alone; K is symmetric, the diagonal is
    (h + 1) * n_i, and an intersection never exceeds either self similarity."""
    X = gen(10000, 40, 0)
    e = WeisfeilerLehmanOptimalAssignment(n_iter=5)
    K = e.fit_transform(X)
    # vertices of a graph = keys of its edge dictionary: ~1.5 % of the ER vertices have no edge and drop out
    n = np.diff(pack(X, "wloa", len_ok=lambda k: k >= 2).graph_ptr).astype(float)
    assert n.sum() < sum(len(l) for _, l in X)
    _eq(np.diagonal(K), 6 * n)
    assert np.array_equal(K, K.T)
    assert np.all(K <= np.minimum.outer(6 * n, 6 * n))
    pick = np.random.RandomState(3).choice(len(X), 48, replace=False)
    _eq(K[np.ix_(pick, pick)], WLOAOracle(n_iter=5).fit_transform([X[i] for i in pick]))
    assert e.stats_.n_entries == 6 * int(n.sum())  # one unary entry per (vertex, level)


# ------------------------------------------------------------------ two independent restatements agree
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_wloa_oracle_equals_histogram_intersection_of_the_wl_oracle(seed):
    """WL-OA = histogram intersection of the WL-subtree feature vectors of all levels, on the vertices that touch an
    edge.  Rebuild it from the (separately pinned) WL oracle's per-level labels and compare with WLOAOracle."""
    from collections import Counter
    from oracle.gk_oracle import WLOracle
    rs = np.random.RandomState(seed)
    X = gen(24, 10, 100 + seed, nl=3)
    for g, _l in X:  # thin the graphs: isolated vertices and one-directional edges appear
        for e in [e for e in sorted(g) if rs.rand() < 0.35]:
            del g[e]
    X = [x for x in X if len(x[0])]
    h = 3
    # the WL oracle walks every labelled vertex; restrict it to the edge-dictionary vertices first
    Xr = []
    for g, l in X:
        verts = {x for e in g for x in e}
        Xr.append([g, {v: l[v] for v in verts}])
    _, levels = WLOracle(n_iter=h).fit_transform(Xr, return_levels=True)
    feats = []
    for j in range(len(Xr)):
        c = Counter()
        for lv in range(h + 1):
            c.update((lv, lab) for lab in levels[lv][j].values())
        feats.append(c)
    K = np.array([[sum(min(n, b[k]) for k, n in a.items() if k in b) for b in feats] for a in feats], dtype=float)
    _eq(WLOAOracle(n_iter=h).fit_transform(X), K)
