"""SURVEY 8(f) "next" rows: EdgeHistogram, WeisfeilerLehman over EdgeHistogram / ShortestPath,
CoreFramework.  Goldens come from the REAL reference (tests/golden/make_golden_next.py).

CPU (-m "not gpu"): the oracle restatements and the HOST logic of the new estimators (packing,
level-tagged labels, k-core subgraph blocks, row map) through the numpy model of the device
pipeline (tests/blockref.py).  GPU (-m gpu): the CUDA path through the C-ABI."""
import os
import warnings

import numpy as np
import pytest

import blockref
import gio
from grakel_b200 import CoreFramework, EdgeHistogram, GraphKernel, ShortestPath, VertexHistogram, WeisfeilerLehman
from grakel_b200.core_framework import core_numbers
from grakel_b200.packing import Block, pack
from oracle.gk_oracle import EHOracle, OGraph, WLOracle, gen, gen_edge_labelled

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")
REF = np.load(os.path.join(G, "next_rows.npz"))


def _sets():
    X = gen_edge_labelled(36, 10, 3)
    return X[:26], X[26:]


def _core_sets():
    C = gen(30, 12, 5, as_adj=True)
    rs = np.random.RandomState(7)
    for el in C:
        A = el[0]
        n = A.shape[0]
        extra = np.triu(rs.rand(n, n) < 0.25, 1)
        A[:] = ((A + extra + extra.T) > 0).astype(float)
    return C[:22], C[22:]


def _eq(a, b, exact=True):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    assert a.shape == b.shape
    if exact:
        assert np.array_equal(a, b, equal_nan=True), f"max abs diff {np.nanmax(np.abs(a - b))}"
    else:
        np.testing.assert_allclose(a, b, rtol=1e-12, atol=0, equal_nan=True)


# ------------------------------------------------------------------ oracle vs the real reference
@pytest.mark.parametrize("tag,nrm", [("", False), ("_n", True)])
def test_oracle_matches_reference(tag, nrm):
    fit, new = _sets()
    o = EHOracle(normalize=nrm)
    _eq(o.fit_transform(fit), REF["eh_K" + tag])
    _eq(o.transform(new), REF["eh_Kt" + tag])
    for name, base in (("wleh", "edge_histogram"), ("wlsp", "shortest_path")):
        o = WLOracle(n_iter=2, base=base, normalize=nrm)
        _eq(o.fit_transform(fit), REF[f"{name}_K{tag}"])
        _eq(o.transform(new), REF[f"{name}_Kt{tag}"])


def test_core_numbers_match_the_reference_algorithm():
    """Vectorised peeling == the Batagelj-Zaversnik loop of core_framework.py:378-409 (restated here on
    the oracle's neighbour dictionaries), incl. isolated vertices and a self loop."""
    def bz(ed):
        nbrs = {v: list(d.keys()) for v, d in ed.items()}
        degrees = {v: len(n) for v, n in nbrs.items()}
        nodes = sorted(degrees, key=degrees.get)
        bins, cur = [0], 0
        for i, v in enumerate(nodes):
            if degrees[v] > cur:
                bins.extend([i] * (degrees[v] - cur))
                cur = degrees[v]
        pos = {v: p for p, v in enumerate(nodes)}
        core = degrees
        for v in nodes:
            for u in nbrs[v]:
                if core[u] > core[v]:
                    nbrs[u].remove(v)
                    p, s = pos[u], bins[core[u]]
                    pos[u], pos[nodes[s]] = s, p
                    nodes[s], nodes[p] = nodes[p], nodes[s]
                    bins[core[u]] += 1
                    core[u] -= 1
        return core
    fit, new = _core_sets()
    X = fit + new
    A = np.zeros((6, 6))
    A[0, 1] = A[1, 0] = A[1, 2] = A[2, 1] = A[0, 2] = A[2, 0] = A[3, 3] = 1  # triangle, self loop, two isolated
    X.append([A, {i: 0 for i in range(6)}])
    b = pack(X, "sp", need_labels=True, len_ok=lambda n: n >= 1)
    c = core_numbers(b)
    for g, x in enumerate(X):
        ref = bz(OGraph(x[0], x[1]).ed)
        v0 = int(b.graph_ptr[g])
        assert [int(c[v0 + i]) for i in range(len(ref))] == [ref[i] for i in range(len(ref))]


# ------------------------------------------------------------------ host logic through the block model
def _host_fit_transform(est, fit, new, model):
    est._method_calling = 2
    est.initialize()
    fx = est.parse_input(fit)
    Kx = model(fx, None)
    est.X = fx
    est._nx = fx.block.n_graphs if not hasattr(fx, "n_rows") else fx.n_rows
    if hasattr(est, "_inv_labels") is False and isinstance(est, WeisfeilerLehman):
        pass
    est._method_calling = 3
    fy = est.parse_input(new)
    return Kx, model(fx, fy)


@pytest.mark.parametrize("nrm", [False, True])
def test_host_logic_edge_histogram_and_wl_bases(nrm):
    fit, new = _sets()
    tag = "_n" if nrm else ""

    def vh_model(levels):
        def m(fx, fy):
            if fy is None:
                return blockref.wl_gram_block(fx.block, fx.ids, levels, normalize=nrm)[0]
            return blockref.wl_gram_block(Block.concat(fx.block, fy.block), np.concatenate([fx.ids, fy.ids]), levels,
                                          n_fit=fx.block.n_graphs, normalize=nrm)[0]
        return m

    def sp_model(levels):
        def m(fx, fy):
            if fy is None:
                return blockref.sp_gram_block(fx.block, fx.ids, wl_iter=levels, normalize=nrm)[0]
            return blockref.sp_gram_block(Block.concat(fx.block, fy.block), np.concatenate([fx.ids, fy.ids]),
                                          n_fit=fx.block.n_graphs, wl_iter=levels, normalize=nrm)[0]
        return m

    with np.errstate(all="ignore"):
        Kx, Kt = _host_fit_transform(EdgeHistogram(), fit, new, vh_model(0))
        _eq(Kx, REF["eh_K" + tag]); _eq(Kt, REF["eh_Kt" + tag])
        Kx, Kt = _host_fit_transform(WeisfeilerLehman(n_iter=2, base_graph_kernel=EdgeHistogram), fit, new, vh_model(2))
        _eq(Kx, REF["wleh_K" + tag]); _eq(Kt, REF["wleh_Kt" + tag])
        Kx, Kt = _host_fit_transform(WeisfeilerLehman(n_iter=2, base_graph_kernel=ShortestPath), fit, new, sp_model(2))
        _eq(Kx, REF["wlsp_K" + tag]); _eq(Kt, REF["wlsp_Kt" + tag])


@pytest.mark.parametrize("nrm", [False, True])
def test_host_logic_core_framework(nrm):
    fit, new = _core_sets()
    tag = "_n" if nrm else ""
    for name, base, wl_levels in (("corewl", (WeisfeilerLehman, {"n_iter": 2}), 2), ("coresp", ShortestPath, None)):
        est = CoreFramework(base_graph_kernel=base)
        est._method_calling = 2
        est.initialize()
        fx = est.parse_input(fit)
        est.X = fx
        est._nx = fx.n_rows
        est._method_calling = 3
        fy = est.parse_input(new)
        both = Block.concat(fx.block, fy.block)
        ids = np.concatenate([fx.ids, fy.ids])
        rows = np.concatenate([fx.rows, fy.rows + fx.n_rows])
        with np.errstate(all="ignore"):
            if wl_levels is not None:
                Kx = blockref.wl_gram_block(fx.block, fx.ids, wl_levels, normalize=nrm, rows=fx.rows, n_rows=fx.n_rows)[0]
                Kt = blockref.wl_gram_block(both, ids, wl_levels, n_fit=fx.n_rows, normalize=nrm, rows=rows,
                                            n_rows=fx.n_rows + fy.n_rows)[0]
            else:
                Kx = blockref.sp_gram_block(fx.block, fx.ids, normalize=nrm, rows=fx.rows, n_rows=fx.n_rows,
                                            nan_to_num=True)[0]
                Kt = blockref.sp_gram_block(both, ids, n_fit=fx.n_rows, normalize=nrm, rows=rows,
                                            n_rows=fx.n_rows + fy.n_rows, nan_to_num=True)[0]
        _eq(Kx, REF[f"{name}_K{tag}"])
        _eq(Kt, REF[f"{name}_Kt{tag}"])


def test_error_behaviour_of_the_new_estimators():
    fit, _ = _sets()
    with pytest.raises(TypeError):  # EdgeHistogram needs [graph, node labels, edge labels]  (edge_histogram.py:88-99)
        EdgeHistogram().parse_input([fit[0][:2]])
    with pytest.raises(TypeError):
        EdgeHistogram().parse_input(5)
    with pytest.raises(ValueError):
        e = EdgeHistogram(); e._method_calling = 1; e.parse_input([[]])
    with pytest.raises(TypeError):
        CoreFramework(base_graph_kernel=(5, {})).initialize()
    with pytest.raises(ValueError):
        CoreFramework(base_graph_kernel=(ShortestPath, 5)).initialize()
    c = CoreFramework(min_core=3)
    assert c.min_core == -1  # core_framework.py:50 stores -1 whatever was passed; results must match the reference
    gk = GraphKernel(kernel=[{"name": "core_framework"}, {"name": "WL", "n_iter": 2}, "VH"])
    gk.initialize()
    assert isinstance(gk.kernel_, CoreFramework) and gk.kernel_.base_graph_kernel[0] is WeisfeilerLehman
    gk = GraphKernel(kernel="EH"); gk.initialize()
    assert isinstance(gk.kernel_, EdgeHistogram)


# ------------------------------------------------------------------ the CUDA path
@pytest.mark.gpu
@pytest.mark.parametrize("tag,nrm", [("", False), ("_n", True)])
def test_gpu_edge_histogram_and_wl_bases(tag, nrm):
    fit, new = _sets()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        e = EdgeHistogram(normalize=nrm)
        _eq(e.fit_transform(fit), REF["eh_K" + tag]); _eq(e.transform(new), REF["eh_Kt" + tag])
        for name, cls in (("wleh", EdgeHistogram), ("wlsp", ShortestPath)):
            w = WeisfeilerLehman(n_iter=2, base_graph_kernel=cls, normalize=nrm)
            _eq(w.fit_transform(fit), REF[f"{name}_K{tag}"])
            _eq(w.transform(new), REF[f"{name}_Kt{tag}"])
            if not nrm:
                _eq(w.diagonal()[0], np.diagonal(REF[f"{name}_K"]))


@pytest.mark.gpu
def test_gpu_wl_sp_on_mutag_prefix():
    M = gio.dec_dataset(gio.load(os.path.join(G, "mutag_graphs.json.gz")))[:48]
    _eq(WeisfeilerLehman(n_iter=3, base_graph_kernel=ShortestPath).fit_transform(M), REF["wlsp_mutag_h3"])
    gk = GraphKernel(kernel=[{"name": "WL", "n_iter": 3}, "SP"])
    _eq(gk.fit_transform(M), REF["wlsp_mutag_h3"])


@pytest.mark.gpu
@pytest.mark.parametrize("tag,nrm", [("", False), ("_n", True)])
def test_gpu_core_framework(tag, nrm):
    fit, new = _core_sets()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for name, base in (("corewl", (WeisfeilerLehman, {"n_iter": 2})), ("coresp", ShortestPath)):
            c = CoreFramework(base_graph_kernel=base, normalize=nrm)
            _eq(c.fit_transform(fit), REF[f"{name}_K{tag}"])
            _eq(c.transform(new), REF[f"{name}_Kt{tag}"])
        c = CoreFramework(base_graph_kernel=(WeisfeilerLehman, {"n_iter": 2}))
        c.fit(fit)
        _eq(c.diagonal(), np.diagonal(REF["corewl_K"]))
        _eq(c.transform(new), REF["corewl_Kt"])
        gk = GraphKernel(kernel=[{"name": "CORE"}, {"name": "WL", "n_iter": 2}])
        _eq(gk.fit_transform(fit), REF["corewl_K"])
