"""CPU tests: host packing / label dictionaries / estimator protocol / C-ABI symbols.
No GPU compute here (run with -m "not gpu")."""
import ctypes
import os
import pickle
import re
import warnings

import numpy as np
import pytest

import blockref
import gio
from grakel_b200 import GraphKernel, ShortestPath, VertexHistogram, WeisfeilerLehman, _lib
from grakel_b200.packing import Block, Graph, label_ids, pack

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")


def _wl_host(X, Y, h, norm):
    bx = pack(X, "wl", len_ok=lambda n: n >= 2)
    ix, dic = label_ids(bx.labels, None)
    Kx, _, _ = blockref.wl_gram_block(bx, ix, h, normalize=norm)
    if Y is None:
        return Kx, None
    by = pack(Y, "wl")
    iy, _ = label_ids(by.labels, dic)
    Kt, _, _ = blockref.wl_gram_block(Block.concat(bx, by), np.concatenate([ix, iy]), h, n_fit=bx.n_graphs,
                                      normalize=norm)
    return Kx, Kt


def _sp_host(X, Y, wlab, alg, norm):
    kw = dict(need_labels=wlab, want_weights=True, fw_zero_is_absent=alg == "floyd_warshall",
              len_ok=lambda n: n in (2, 3) or (n == 1 and not wlab))
    bx = pack(X, "sp", **kw)
    ix, dic = label_ids(bx.labels, None, sort_new=False) if wlab else (np.zeros(bx.n_vertices, int), {})
    with np.errstate(all="ignore"):
        Kx, _, _ = blockref.sp_gram_block(bx, ix, with_labels=wlab, normalize=norm)
        if Y is None:
            return Kx, None
        by = pack(Y, "sp", **kw)
        iy = label_ids(by.labels, dic, sort_new=False)[0] if wlab else np.zeros(by.n_vertices, int)
        Kt, _, _ = blockref.sp_gram_block(Block.concat(bx, by), np.concatenate([ix, iy]), n_fit=bx.n_graphs,
                                          with_labels=wlab, normalize=norm)
    return Kx, Kt


def _check(X, Y, out, skip_dijkstra_real=False):
    for key, rec in out.items():
        parts = key.split("_")
        if "error" in rec:
            continue
        if parts[0] == "wl":
            Kx, Kt = _wl_host(X, Y, int(parts[1][1:]), parts[2] == "n")
        else:
            alg = "_".join(parts[2:-1])
            if skip_dijkstra_real and alg == "dijkstra":
                continue
            Kx, Kt = _sp_host(X, Y, parts[1] == "l", alg, parts[-1] == "n")
        assert np.array_equal(Kx, np.asarray(rec["fit_transform"], dtype=float), equal_nan=True), key
        if "transform" in rec:
            assert np.array_equal(Kt, np.asarray(rec["transform"], dtype=float), equal_nan=True), key


def test_packing_all_spellings_match_reference():
    g = gio.load(os.path.join(G, "spellings.json.gz"))
    for name, case in g["cases"].items():
        _check(gio.dec_dataset(case["X"]), None, case["out"])


@pytest.mark.parametrize("tag", ["unit", "intw"])
def test_joint_relabel_equals_fit_dictionaries(tag):
    """transform(Y) through joint relabelling of X||Y == the reference's dictionary replay,
    on data whose test split contains labels unseen at fit time."""
    d = gio.load(os.path.join(G, "fit_transform.json.gz"))[tag]
    _check(gio.dec_dataset(d["X"]), gio.dec_dataset(d["Y"]), d["out"])


def test_mutag_host_model():
    X = gio.dec_dataset(gio.load(os.path.join(G, "mutag_graphs.json.gz")))
    ref = np.load(os.path.join(G, "mutag_out.npz"))
    K, _ = _wl_host(X, None, 3, False)
    assert np.array_equal(K, ref["wl_h3"])


def test_block_concat_offsets():
    X = gio.dec_dataset(gio.load(os.path.join(G, "spellings.json.gz"))["cases"]["mixed_isolated"]["X"])
    b = pack(X, "wl", len_ok=lambda n: n >= 2)
    c = Block.concat(b, b)
    assert c.n_graphs == 2 * b.n_graphs and c.n_vertices == 2 * b.n_vertices
    assert np.array_equal(c.col_idx[len(b.col_idx):], b.col_idx + b.n_vertices)
    assert c.row_ptr[-1] == 2 * len(b.col_idx)
    for g in range(c.n_graphs):  # neighbours stay inside their graph
        v0, v1 = c.graph_ptr[g], c.graph_ptr[g + 1]
        nb = c.col_idx[c.row_ptr[v0]:c.row_ptr[v1]]
        assert nb.size == 0 or (nb.min() >= v0 and nb.max() < v1)


def test_error_behaviour_matches_reference():
    wl = WeisfeilerLehman(n_iter=2)
    with pytest.raises(TypeError):
        wl.fit_transform(5)  # weisfeiler_lehman.py:144
    with pytest.raises(ValueError):
        wl.fit_transform(None)  # :319
    with pytest.raises(ValueError):
        wl.fit([])  # :194 parsed input is empty
    with pytest.raises(TypeError):
        WeisfeilerLehman(n_iter=0).fit([[{(0, 1): 1}, {0: 1, 1: 2}]])  # :112-113
    with pytest.raises(TypeError):
        WeisfeilerLehman(n_iter=2, base_graph_kernel=3).fit([[{(0, 1): 1}, {0: 1, 1: 2}]])  # :87
    with pytest.raises(TypeError):
        wl.fit([[{(0, 1): 1}]])  # element of length 1 (:182)
    with pytest.raises(ValueError):
        ShortestPath(algorithm_type="bfs").fit([[{(0, 1): 1}, {0: 1, 1: 2}]])  # shortest_path.py:252
    with pytest.raises(ValueError):
        wl.fit([["not a graph", {0: 1}]])  # graph.py:214
    with pytest.raises(TypeError):
        wl.fit([[{(0, 1): 1, (1, 0): 1}, {0: "a", 1: 3}]])  # unsortable label alphabet (:204)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        wl.fit([[], [{(0, 1): 1, (1, 0): 1}, {0: 1, 1: 2}]])  # empty element is skipped (:153-155)
        assert any("Ignoring empty element" in str(x.message) for x in w)
    assert wl._nx == 1
    from sklearn.exceptions import NotFittedError
    with pytest.raises(NotFittedError):
        WeisfeilerLehman().transform([[{(0, 1): 1}, {0: 1, 1: 2}]])


def test_fitted_estimators_pickle_and_clone():
    """grakel/tests/test_common.py:53-58: fitted kernels must pickle; sklearn clone / set_params work."""
    from sklearn.base import clone
    X = gio.dec_dataset(gio.load(os.path.join(G, "spellings.json.gz"))["cases"]["dict_tuple"]["X"])
    for est in (WeisfeilerLehman(n_iter=2, normalize=True), ShortestPath(), VertexHistogram(),
                GraphKernel(kernel=[{"name": "WL", "n_iter": 2}, "subtree_wl"])):
        est.fit(X)
        e2 = pickle.loads(pickle.dumps(est))
        assert type(e2) is type(est)
        c = clone(est)
        assert c.get_params().keys() == est.get_params().keys()
    wl = WeisfeilerLehman(n_iter=2).fit(X)
    assert wl._n_iter == 3 and wl._nx == 2 and 0 in wl._inv_labels
    wl.set_params(n_iter=4)
    assert wl._initialized["n_iter"] is False
    wl.fit(X)
    assert wl._n_iter == 5


def test_graphkernel_dispatch():
    gk = GraphKernel(kernel=[{"name": "weisfeiler_lehman", "n_iter": 3}, {"name": "vertex_histogram"}], normalize=True)
    gk.initialize()
    assert isinstance(gk.kernel_, WeisfeilerLehman) and gk.kernel_.n_iter == 3 and gk.kernel_.normalize is True
    gk = GraphKernel(kernel={"name": "SP", "with_labels": False})
    gk.initialize()
    assert isinstance(gk.kernel_, ShortestPath) and gk.kernel_.with_labels is False
    gk = GraphKernel(kernel="ST-WL")
    gk.initialize()
    assert isinstance(gk.kernel_, VertexHistogram)
    with pytest.raises(ValueError):
        GraphKernel(kernel="no_such_kernel").initialize()
    with pytest.raises(NotImplementedError):
        GraphKernel(kernel="random_walk").initialize()


def test_graph_carrier_accepted():
    A = np.array([[0, 1], [1, 0]])
    b = pack([Graph(A, {0: "a", 1: "b"})], "wl", len_ok=lambda n: n >= 2)
    assert b.n_graphs == 1 and b.labels == ["a", "b"]


def test_cabi_library_exports_every_declared_symbol():
    """The shared library must load on a CPU-only box and export every gk_* entry point
    that include/grakel_b200.h declares (no compute calls here)."""
    hdr = open(os.path.join(ROOT, "include", "grakel_b200.h")).read()
    declared = set(re.findall(r"\b(gk_[a-z0-9_]+)\s*\(", hdr))
    assert {"gk_create", "gk_pack_csr", "gk_wl_features", "gk_sp_features", "gk_gram",
            "gk_wl_fit_transform"} <= declared
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert declared == set(_lib.exported_symbols())
    _lib.load_library().gk_version.restype = ctypes.c_int
    assert _lib.load_library().gk_version() >= 100


def test_no_gpu_means_loud_failure():
    """Without a B200 the engine must raise, never fall back to a CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(Exception):
        WeisfeilerLehman(n_iter=1).fit_transform([[{(0, 1): 1, (1, 0): 1}, {0: 1, 1: 2}]])


# ------------------------------------------------------------------ vectorised packer fast path
def _lab_list(blk):
    """labels as a list of Python objects (the C packer returns an int64 array when every label is an exact int)"""
    return [] if blk.labels is None else [int(x) if isinstance(x, np.integer) else x for x in list(blk.labels)]


def _blocks_equal(a, b, same_order=True):
    if same_order:
        return (np.array_equal(a.graph_ptr, b.graph_ptr) and np.array_equal(a.row_ptr, b.row_ptr)
                and np.array_equal(a.col_idx, b.col_idx) and _lab_list(a) == _lab_list(b)
                and ((a.weights is None) == (b.weights is None))
                and (a.weights is None or np.array_equal(a.weights, b.weights)))
    # vertex order inside a graph may differ: compare the labelled edge multisets per graph
    if not np.array_equal(a.graph_ptr, b.graph_ptr):
        return False
    for blk in (a, b):
        blk._canon = []
        for g in range(blk.n_graphs):
            v0, v1 = blk.graph_ptr[g], blk.graph_ptr[g + 1]
            lab = _lab_list(blk)
            es = sorted((lab[v], lab[blk.col_idx[k]]) for v in range(v0, v1)
                        for k in range(blk.row_ptr[v], blk.row_ptr[v + 1]))
            blk._canon.append((sorted(lab[v0:v1]), es))
    return a._canon == b._canon


class _BlockrefEngine:
    """numpy stand-in for the device engine (tests/blockref.py) behind the estimators' _run"""

    def __init__(self):
        self._lock = __import__("threading").RLock()

    def pack(self, gp, rp, ci, ids, weights, attrs):
        from grakel_b200.packing import Block
        self.block, self.ids = Block(gp, rp, ci, weights, None, attrs), ids

    def wl_features(self, n_iter):
        self.h = n_iter
        from grakel_b200._lib import GkStats
        return GkStats()

    def gram(self, n_graphs, n_fit=None, normalize=False, nan_to_num=False, out=None, stats=None, **kw):
        K, xd, yd = blockref.wl_gram_block(self.block, self.ids, self.h, n_fit=n_fit, normalize=normalize)
        return (None if out is False else K), xd, (yd if n_fit != n_graphs else None)


@pytest.mark.parametrize("kind", ["VH", "EH"])
def test_histogram_kernels_fit_then_normalized_transform(kind, monkeypatch):
    """kernel.py:160-164: `fit(X); transform(Y)` with normalize=True needs X's diagonal although fit_transform
    never ran -- the joint X||Y run returns it (AttributeError on _X_diag before)."""
    from contextlib import contextmanager
    from grakel_b200 import EdgeHistogram, VertexHistogram, _lib
    from oracle.gk_oracle import gen
    eng = _BlockrefEngine()

    @contextmanager
    def fake_engine(device=None):
        yield eng
    monkeypatch.setattr(_lib, "engine", fake_engine)
    X, Y = gen(12, 10, 1), gen(5, 10, 2)
    if kind == "EH":
        add = lambda Z: [[g, L, {e: (L[e[0]] + L[e[1]]) % 3 for e in g}] for g, L in Z]
        X, Y = add(X), add(Y)
        est = EdgeHistogram(normalize=True)
    else:
        est = VertexHistogram(normalize=True)
    K = est.fit(X).transform(Y)
    assert K.shape == (5, 12) and np.all(np.isfinite(K)) and K.max() <= 1.0 + 1e-12
    xd, yd = est.diagonal()
    assert len(xd) == 12 and len(yd) == 5
    Kfull = type(est)(normalize=True).fit_transform(X + Y)
    assert np.allclose(Kfull[12:, :12], K)


def test_c_packer_threads_labels_and_hubs():
    """csrc/fastpack.c: any thread count builds the same block; non-int labels come back as the label objects;
    a hub vertex with unsorted neighbours gets a sorted row (two counting sorts, no quadratic insertion sort)."""
    from grakel_b200 import packing
    from oracle.gk_oracle import gen
    cmod = packing._fastpack
    assert cmod is not None
    X = gen(400, 30, 11)
    for mode_i in (0, 1):
        ref = cmod.pack_edge_dicts(X, mode_i, 1, 1, 1)
        for nt in (2, 3, 8, 64):
            out = cmod.pack_edge_dicts(X, mode_i, 1, 1, nt)
            assert all(bytes(a) == bytes(b) for a, b in zip(ref[:3], out[:3])) and bytes(ref[4]) == bytes(out[4])
        assert isinstance(ref[4], bytearray) and ref[3] is None and ref[5] == 0
    # string labels on some vertices: label objects, in vertex order, identical to the general path
    Xs = [[g, {v: ("c%d" % l if v % 3 == 0 else l) for v, l in L.items()}] for g, L in X]
    for mode in ("wl", "sp"):
        fast = packing.pack(Xs, mode, len_ok=lambda n: n >= 2)
        assert isinstance(fast.labels, list)
        saved, packing._fastpack = packing._fastpack, None
        try:
            slow = packing.pack(Xs, mode, len_ok=lambda n: n >= 2)
        finally:
            packing._fastpack = saved
        assert _blocks_equal(fast, slow)
    # a star with 5 000 leaves listed in descending order, weights carried along
    n = 5000
    g = {}
    for v in range(n, 0, -1):
        g[(0, v)] = float(v)
        g[(v, 0)] = 1.0
    hub = [[g, {v: v % 5 for v in range(n + 1)}]] * 3
    blk = packing.pack(hub, "wl", len_ok=lambda n: n >= 2, want_weights=True)
    for k in range(3):
        v0 = k * (n + 1)
        row = blk.col_idx[blk.row_ptr[v0]:blk.row_ptr[v0 + 1]] - v0
        assert np.array_equal(row, np.arange(1, n + 1))
        assert np.array_equal(blk.weights[blk.row_ptr[v0]:blk.row_ptr[v0 + 1]], np.arange(1, n + 1, dtype=float))


def test_label_ids_integer_arrays_match_the_object_path():
    from grakel_b200.packing import label_ids
    rs = np.random.RandomState(4)
    for labels in (rs.randint(-5, 40, size=3000), rs.randint(0, 2 ** 40, size=500), np.array([7], dtype=np.int64)):
        for sort_new in (True, False):
            ids_a, new_a = label_ids(labels.astype(np.int64), None, sort_new)
            ids_o, new_o = label_ids([int(x) for x in labels], None, sort_new)
            assert np.array_equal(ids_a, ids_o) and new_a == new_o and list(new_a) == list(new_o)
            known = {int(x): i for i, x in enumerate(sorted(set(labels.tolist()))[::2])}
            ids_a, new_a = label_ids(labels.astype(np.int64), known, sort_new)
            ids_o, new_o = label_ids([int(x) for x in labels], known, sort_new)
            assert np.array_equal(ids_a, ids_o) and new_a == new_o


def test_host_delivery_bands_widen_and_mirror():
    """host_deliver.h without a device: fp32 -> float64 through the band / widen / 8x8-transpose-mirror code of
    gk_gram's result delivery, for sizes around every blocking boundary, plus the fp64 normalisation."""
    from grakel_b200 import _lib
    lib = _lib.load_library()

    class Host:  # selftest_deliver needs no handle
        pass
    e = Host()
    e.lib = lib
    e._check = lambda rc: (_ for _ in ()).throw(RuntimeError(lib.gk_last_error().decode())) if rc else None
    deliver = lambda *a, **k: _lib.Engine.selftest_deliver(e, *a, **k)
    rs = np.random.RandomState(0)
    for n in (1, 7, 8, 9, 33, 100, 257, 1000, 2049):
        A = rs.randint(0, 1 << 24, size=(n, n)).astype(np.float32)
        A = np.triu(A) + np.triu(A, 1).T
        assert np.array_equal(deliver(A, 0), A.astype(np.float64)), n
        assert np.array_equal(deliver(A, 1), A.astype(np.float64)), n
        A16 = np.floor(A / 256.0).astype(np.float32)  # integers < 2^16: the two-byte transport
        assert np.array_equal(deliver(A16, 3), A16.astype(np.float64)), n
        d = rs.randint(0, 50, size=n).astype(np.float64)
        with np.errstate(all="ignore"):
            ref = A.astype(np.float64) / np.sqrt(np.outer(d, d))  # kernel.py:198-203
        assert np.array_equal(deliver(A, 2, diag=d), ref, equal_nan=True), n
        assert np.array_equal(deliver(A, 2, diag=d, nan_to_num=True), np.nan_to_num(ref)), n
    R = rs.rand(300, 777).astype(np.float32)
    assert np.array_equal(deliver(R, 1), R.astype(np.float64))
    # pooled result buffers: fresh, writable, C-order; a freed block is handed out again
    import gc
    K = _lib.host_matrix(2000, 2000)
    assert K.flags.c_contiguous and K.flags.writeable and K.dtype == np.float64 and K.shape == (2000, 2000)
    K[:] = 1.0
    addr = K.ctypes.data
    del K
    gc.collect()
    assert _lib.host_matrix(2000, 2000).ctypes.data == addr


@pytest.mark.parametrize("mode", ["wl", "sp", "wloa"])
def test_fast_edge_dict_path_equals_the_general_packer(mode, monkeypatch):
    """{(u, v): w} inputs with integer symbols take a vectorised path (packing._fast_edge_dict); it must build the
    block the general path builds, and hand every unusual element back to it."""
    from grakel_b200 import packing
    from oracle.gk_oracle import gen
    rs = np.random.RandomState(3)
    X = gen(60, 14, 5)
    for i, (g, l) in enumerate(X):  # weights, a shifted id range, directed edges, dropped edges
        if i % 4 == 1:
            for e in list(g):
                g[e] = float(rs.randint(1, 4))
        if i % 4 == 2:
            X[i][0] = {(a + 7, b + 7): w for (a, b), w in g.items()}
            X[i][1] = {v + 7: lab for v, lab in l.items()}
        if i % 4 == 3:
            for e in [e for e in sorted(g) if rs.rand() < 0.3]:
                del g[e]
    X = [x for x in X if len(x[0])]
    odd = [
        [{("a", "b"): 1, ("b", "a"): 1}, {"a": 0, "b": 1}],           # symbols that are not integers
        [{(0, 1): 1, (1, 0): 1}, {0: 5, 1: 6, 3: 7}],                 # label keys with a gap
        [{(0, 1): 1, (1, 2): 1}, {1: 5, 0: 6, 2: 7}],                 # label keys out of order
        [{(0, 1): 1.0, (5, 0): 1.0, (1, 0): 2.0}, {0: 1, 1: 2, 5: 3}],
    ]
    kw = dict(len_ok=lambda n: n >= 2, want_weights=True)
    cmod = packing._fastpack
    assert cmod is not None, "grakel_b200/_fastpack was not built (grakel_b200/csrc/build.sh)"
    # the CPython-level packer (csrc/fastpack.c) takes the regular elements in one call ...
    cres = cmod.pack_edge_dicts(X, 0 if mode == "wl" else 1, 1)
    assert cres is not None
    cblock = packing.pack(X, mode, **kw)
    assert cmod.pack_edge_dicts(X + odd, 0 if mode == "wl" else 1, 1) is None  # ... and declines the odd ones
    monkeypatch.setattr(packing, "_fastpack", None)
    assert _blocks_equal(cblock, packing.pack(X, mode, **kw))  # == the numpy fast path, bit for bit
    fast = packing.pack(X + odd, mode, **kw)
    calls = []
    real = packing._fast_edge_dict
    monkeypatch.setattr(packing, "_fast_edge_dict", lambda *a: calls.append(1) or None)
    slow = packing.pack(X + odd, mode, **kw)
    assert len(calls) == len(X) + len(odd)
    assert _blocks_equal(fast, slow, same_order=mode != "wloa")
    assert _blocks_equal(cblock, packing.pack(X, mode, **kw), same_order=mode != "wloa")  # == the general path
    monkeypatch.setattr(packing, "_fast_edge_dict", real)
    taken = sum(real(g, l, mode, True) is not None for g, l in X)
    assert taken == len(X)  # the regular elements really go through the fast path
    assert real(odd[0][0], odd[0][1], mode, True) is None
    # an unlabelled vertex: the fast path declines, the general path raises / skips exactly as before
    bad = [[{(0, 1): 1, (1, 0): 1}, {0: 1}]]
    assert real(bad[0][0], bad[0][1], mode, True) is None
    assert cmod.pack_edge_dicts(bad, 0 if mode == "wl" else 1, 1) is None
    monkeypatch.setattr(packing, "_fastpack", cmod)
    with pytest.raises(KeyError):
        packing.pack(bad, mode, **kw)
    # inputs the C packer must leave alone: floats as symbols, bool keys, a non-dict graph, huge ints, empty graph
    for weird in ([[{(0.0, 1.0): 1}, {0.0: 1, 1.0: 2}]], [[{(0, 1): "x"}, {0: 1, 1: 2}]], [[[(0, 1)], {0: 1, 1: 2}]],
                  [[{(0, 2 ** 70): 1}, {0: 1, 2 ** 70: 2}]], [[{}, {0: 1}]], [[{(0, 1, 2): 1}, {0: 1, 1: 2}]]):
        assert cmod.pack_edge_dicts(weird, 0 if mode == "wl" else 1, 1) is None


def _golden_inv_labels(name):
    """(generator parameters, {level: dictionary}) of one case of tests/golden/inv_labels.json.gz."""
    import gzip
    import json
    with gzip.open(os.path.join(G, "inv_labels.json.gz"), "rt") as f:
        case = json.load(f)[name]
    return case["params"], {int(i): {(k if int(i) else int(k)): v for k, v in pairs}
                            for i, pairs in case["levels"].items()}


@pytest.mark.parametrize("name", ["er60", "er25_deep"])
def test_level_dictionaries_from_a_partition_are_the_references(name):
    """`_inv_labels[i]`, i >= 1 (weisfeiler_lehman.py:223-257): the host half of the export -- class arrays (here the
    oracle's; on the device gk_wl_labels) -> credential strings in the reference's numbering -- against the
    dictionaries the real reference built (tests/golden/make_golden_invlabels.py)."""
    from grakel_b200.kernels import wl_reference_dictionaries
    from oracle.gk_oracle import WLOracle, gen, wl_partitions
    c, want = _golden_inv_labels(name)
    X = gen(c["N"], c["nbar"], c["seed"], nl=c["nl"])
    block = pack(X, "wl", len_ok=lambda n: n >= 2)
    ids, d0 = label_ids(block.labels, None, sort_new=True)
    assert d0 == want[0]
    _, levels = WLOracle(n_iter=c["n_iter"]).fit_transform(X, return_levels=True)
    parts = wl_partitions(levels)
    got, labels = wl_reference_dictionaries(np.asarray(block.row_ptr), np.asarray(block.col_idx), ids, len(d0), parts)
    assert sorted(got) == list(range(1, c["n_iter"] + 1))
    for i in got:
        assert got[i] == want[i], f"level {i}"
    assert len(labels) == c["n_iter"] + 1 and all(len(l) == block.n_vertices for l in labels)


def test_level_dictionaries_are_lazy_and_pickle(monkeypatch):
    X = gio.dec_dataset(gio.load(os.path.join(G, "spellings.json.gz"))["cases"]["dict_tuple"]["X"])
    wl = WeisfeilerLehman(n_iter=2).fit(X)
    calls = []
    monkeypatch.setattr(WeisfeilerLehman, "_export_level_dictionaries",
                        lambda self: calls.append(1) or {1: {"a": 7}, 2: {"b": 9}})
    assert list(wl._inv_labels) == [0] and not calls
    w2 = pickle.loads(pickle.dumps(wl))
    assert list(w2._inv_labels) == [0] and w2._inv_labels._owner is w2
    assert wl._inv_labels[2] == {"b": 9} and wl._inv_labels[1] == {"a": 7} and len(calls) == 1
    with pytest.raises(KeyError):
        wl._inv_labels[3]
    with pytest.raises(KeyError):
        wl._inv_labels["x"]


def _sp_matrix(adj):
    """APSP of a small unit-weight graph (test-local Floyd-Warshall)."""
    n = len(adj)
    S = np.where(np.asarray(adj) > 0, 1.0, np.inf)
    np.fill_diagonal(S, 0.0)
    for k_ in range(n):
        S = np.minimum(S, S[:, [k_]] + S[[k_], :])
    return S


def test_spattr_user_metric_pairwise_operation_and_generic_driver(monkeypatch):
    """ShortestPathAttr with a user metric: `pairwise_operation` (vectorised over path lengths) equals the literal
    4-deep loop of shortest_path.py:130-164, and the generic driver (kernel.py:236-296) fills, mirrors and normalises
    like the reference.  The device side (the shortest-path matrices) is replaced by a host stand-in here; the GPU
    test pins the whole route against real-reference goldens."""
    from grakel_b200 import ShortestPathAttr
    sys_path_golden = os.path.join(os.path.dirname(__file__), "golden")
    import importlib.util
    spec = importlib.util.spec_from_file_location("mk_spattr_metric", os.path.join(sys_path_golden, "make_golden_spattr_metric.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    X = mk.gen_attr(9, 7, 11)
    items = [(_sp_matrix(a), np.asarray([L[i] for i in range(len(a))], dtype=float)) for a, L in X]

    def literal(x, y, metric):
        (Sx, px), (Sy, py) = x, y
        tot = 0
        for i in range(len(Sx)):
            for j in range(len(Sx)):
                if i == j:
                    continue
                for k_ in range(len(Sy)):
                    for m in range(len(Sy)):
                        if k_ == m:
                            continue
                        if Sx[i, j] == Sy[k_, m] and Sx[i, j] != float("Inf"):
                            tot += metric(px[i], py[k_]) * metric(px[j], py[m])
        return tot

    est = ShortestPathAttr(metric=mk.rbf)
    for a in range(4):
        for b in range(a, 5):
            assert est.pairwise_operation(items[a], items[b]) == pytest.approx(literal(items[a], items[b], mk.rbf), rel=1e-12)
    # the whole estimator route with the device step mocked: goldens of the REAL reference
    ref = np.load(os.path.join(sys_path_golden, "spattr_metric.npz"))
    lookup = {id(x): it for x, it in zip(X, items)}
    monkeypatch.setattr(ShortestPathAttr, "_parse_pairs", lambda self, Xs: [lookup[id(x)] for x in Xs])
    fit, new = X[:6], X[6:]
    e = ShortestPathAttr(metric=mk.rbf)
    K = e.fit_transform(fit)
    np.testing.assert_allclose(K, ref["unit_K"], rtol=1e-10)
    assert np.array_equal(K, K.T)
    np.testing.assert_allclose(e.transform(new), ref["unit_Kt"], rtol=1e-10)
    xd, yd = e.diagonal()
    np.testing.assert_allclose(xd, np.diagonal(ref["unit_K"]), rtol=1e-10)
    assert yd.shape == (3,)
    e = ShortestPathAttr(metric=mk.rbf, normalize=True)
    np.testing.assert_allclose(e.fit_transform(fit), ref["unit_Kn"], rtol=1e-10)
    np.testing.assert_allclose(e.transform(new), ref["unit_Ktn"], rtol=1e-10)
    with pytest.raises(TypeError):
        ShortestPathAttr(metric=3).fit(fit)
