"""CoreFramework (grakel/kernels/core_framework.py:23-400) on the device engine.

The reference computes, for every core level i = max_core .. min_core+1, the base kernel on the
i-core subgraphs of the graphs that have one, and scatter-adds the per-level matrices
(core_framework.py:177-223).  Here all subgraphs of all levels are packed as separate graphs of ONE
block; vertex labels are tagged with their core level, so that no feature column is shared between
levels (each reference level has its own base-kernel instance, i.e. its own dictionaries); the C-ABI
row map (gk_set_row_map) sends every subgraph back to its graph's row, and one Gram over the joint
feature block equals the reference's sum over levels.  Base kernels: ShortestPath (the reference
default), WeisfeilerLehman (over VertexHistogram / ShortestPath), VertexHistogram.
"""
from __future__ import annotations

import warnings
from collections.abc import Iterable

import numpy as np
from sklearn.utils.validation import check_is_fitted

from . import _lib
from .kernels import Kernel, ShortestPath, VertexHistogram, WeisfeilerLehman
from .packing import Block, label_ids, pack


def core_numbers(block):
    """Core number of every vertex of a packed block (Batagelj-Zaversnik result of
    core_framework.py:378-409, computed by vectorised peeling).  A self loop counts as one neighbour
    that is never removed before its vertex, as in the reference's neighbour lists."""
    V = block.n_vertices
    rp = block.row_ptr.astype(np.int64)
    ci = block.col_idx.astype(np.int64)
    deg = np.diff(rp)
    src = np.repeat(np.arange(V, dtype=np.int64), deg)
    cur = deg.copy()
    core = np.zeros(V, dtype=np.int64)
    alive = np.ones(V, dtype=bool)
    k, remaining = 0, V
    while remaining:
        rem = alive & (cur <= k)
        n_rem = int(rem.sum())
        if n_rem == 0:
            k = max(k + 1, int(cur[alive].min()))
            continue
        core[rem] = k
        alive[rem] = False
        remaining -= n_rem
        m = rem[src] & alive[ci]
        if m.any():
            np.subtract.at(cur, ci[m], 1)
    return core


def core_subgraph_block(block, core, tag_labels=True):
    """All non-empty i-core subgraphs (i = max core .. 0) of a block as one block.

    Returns (virtual block, row index of every virtual graph, max core number).  Labels become
    (level, label) pairs; weights and attributes are carried over."""
    N, V = block.n_graphs, block.n_vertices
    gp = block.graph_ptr.astype(np.int64)
    rp = block.row_ptr.astype(np.int64)
    ci = block.col_idx.astype(np.int64)
    deg = np.diff(rp)
    src = np.repeat(np.arange(V, dtype=np.int64), deg)
    vgraph = np.repeat(np.arange(N, dtype=np.int64), np.diff(gp))
    max_core = int(core.max()) if V else 0
    g_ptr, r_ptr, cols, w_parts, labels, rows = [np.zeros(1, dtype=np.int64)], [np.zeros(1, dtype=np.int64)], [], [], [], []
    v_off = e_off = 0
    for level in range(max_core, -1, -1):
        keep = core >= level
        newid = np.cumsum(keep) - 1
        em = keep[src] & keep[ci]
        nv = int(keep.sum())
        sizes = np.bincount(vgraph[keep], minlength=N)
        present = sizes > 0
        g_ptr.append(v_off + np.cumsum(sizes[present]))
        rows.append(np.nonzero(present)[0])
        r_ptr.append(e_off + np.cumsum(np.bincount(newid[src[em]], minlength=nv)))
        cols.append(newid[ci[em]] + v_off)
        if block.weights is not None:
            w_parts.append(block.weights[em])
        if block.labels is not None:
            kept = np.nonzero(keep)[0]
            labels.extend(((level, block.labels[v]) if tag_labels else block.labels[v]) for v in kept.tolist())
        v_off += nv
        e_off += int(em.sum())
    vb = Block(np.concatenate(g_ptr), np.concatenate(r_ptr), np.concatenate(cols) if cols else np.zeros(0, dtype=np.int64),
               np.concatenate(w_parts) if block.weights is not None else None,
               labels if block.labels is not None else None, None, block.all_adjacency)
    return vb, np.concatenate(rows).astype(np.int32), max_core


class _FittedCores:
    def __init__(self, vblock, ids, rows, dictionary, n_rows, max_core):
        self.block, self.ids, self.rows, self.dictionary, self.n_rows, self.max_core = vblock, ids, rows, dictionary, n_rows, max_core


class CoreFramework(Kernel):
    """The core kernel framework (core_framework.py:23-400)."""

    _graph_format = "dictionary"
    _nan_to_num = True

    def __init__(self, n_jobs=None, verbose=False, normalize=False, min_core=-1, base_graph_kernel=None):
        super().__init__(n_jobs=n_jobs, verbose=verbose, normalize=normalize)
        # core_framework.py:50: the reference stores -1 whatever the argument says; results must match
        # the reference's, so the same is done here (sklearn clone/get_params then agree as well).
        self.min_core = -1
        self.base_graph_kernel = base_graph_kernel
        self._initialized.update({"min_core": False, "base_graph_kernel": False})

    def initialize(self):
        super().initialize()
        if not self._initialized["base_graph_kernel"]:  # core_framework.py:59-88
            base = self.base_graph_kernel
            if base is None:
                base, params = ShortestPath, dict()
            elif type(base) is type and issubclass(base, Kernel):
                params = dict()
            else:
                try:
                    base, params = base
                except Exception:
                    raise TypeError("Base kernel was not formulated in the correct way. Check documentation.")
                if not (type(base) is type and issubclass(base, Kernel)):
                    raise TypeError("The first argument must be a valid grakel.kernel.kernel Object")
                if type(params) is not dict:
                    raise ValueError("If the second argument of base kernel exists, it must be a dictionary between "
                                     "parameters names and values")
                params = dict(params)
                params.pop("normalize", None)
            if base not in (ShortestPath, WeisfeilerLehman, VertexHistogram):
                raise NotImplementedError("grakel_b200 runs CoreFramework over ShortestPath, WeisfeilerLehman or "
                                          "VertexHistogram; other base kernels are outside the device hot path")
            params["normalize"] = False
            params["verbose"] = self.verbose
            params["n_jobs"] = None
            self.base_graph_kernel_ = base
            self.params_ = params
            self._base = base(**params)  # validates the parameters exactly like the reference's per-level instances
            self._base.initialize()
            self._initialized["base_graph_kernel"] = True
        if not self._initialized["min_core"]:
            if type(self.min_core) is not int or self.min_core < -1:
                raise TypeError("'min_core' must be an integer bigger than -1")
            self._initialized["min_core"] = True

    # ---- host side ------------------------------------------------------------
    def _pack(self, X):
        if not isinstance(X, Iterable):
            raise TypeError("input must be an iterable\n")
        base = self.base_graph_kernel_
        need_labels = not (base is ShortestPath and not self.params_.get("with_labels", True))
        # every element becomes Graph(x[0], x[1], x[2], "adjacency") first (core_framework.py:133-139):
        # the vertex set is the graph's own (matrix order / sorted symbols that occur in an edge)
        block = pack(X, "sp", need_labels=need_labels, len_ok=lambda n: n >= 1, want_weights=base is not VertexHistogram)
        if block.weights is not None and np.any(block.weights != np.rint(block.weights)):
            raise NotImplementedError("non-integer edge weights are supported with Floyd-Warshall semantics only")
        core = core_numbers(block)
        return core_subgraph_block(block, core) + (block.n_graphs,)

    def parse_input(self, X):
        vb, rows, max_core, n = self._pack(X)
        if max_core <= self.min_core:
            raise ValueError("The maximum core equals the min_core boundary set in init.")
        sort_new = self.base_graph_kernel_ is WeisfeilerLehman
        if vb.labels is None:
            ids, dictionary = None, {}
        elif self._method_calling in (1, 2):
            ids, dictionary = label_ids(vb.labels, None, sort_new=sort_new)
        else:
            ids, _ = label_ids(vb.labels, self.X.dictionary, sort_new=sort_new)
            dictionary = self.X.dictionary
        return _FittedCores(vb, ids, rows, dictionary, n, max_core)

    # ---- device side ----------------------------------------------------------
    def _device_features(self, eng):
        base, b = self.base_graph_kernel_, self._base
        if base is ShortestPath:
            return eng.sp_features(with_labels=bool(b.with_labels))
        if base is VertexHistogram:
            return eng.wl_features(0)
        if b._base_graph_kernel is ShortestPath:
            return eng.wl_sp_features(b._n_iter - 1)
        if b._base_graph_kernel is not VertexHistogram:
            raise NotImplementedError("CoreFramework over WeisfeilerLehman needs the subtree or ShortestPath base kernel")
        return eng.wl_features(b._n_iter - 1)

    def _run_cores(self, block, ids, rows, n_rows, n_fit, want_matrix=True):
        with _lib.engine(getattr(self, "device_", None)) as eng:
            eng.pack(block.graph_ptr, block.row_ptr, block.col_idx, ids, block.weights, block.attrs)
            eng.set_row_map(n_rows, rows)
            self.stats_ = self._device_features(eng)
            K, xd, yd = eng.gram(n_rows, n_fit=n_fit, normalize=bool(self.normalize) and want_matrix,
                                 nan_to_num=True, out=None if want_matrix else False, stats=self.stats_)
        return K, xd, yd

    def fit(self, X, y=None):
        self._is_transformed = False
        self._method_calling = 1
        self.initialize()
        if X is None:
            raise ValueError("`fit` input cannot be None")
        self.X = self.parse_input(X)
        self._nx = self.X.n_rows
        self._max_core_number = self.X.max_core
        if hasattr(self, "_X_diag"):
            delattr(self, "_X_diag")
        return self

    def fit_transform(self, X, y=None):
        self._method_calling = 2
        self._is_transformed = False
        self.initialize()
        if X is None:
            raise ValueError("transform input cannot be None")
        self.X = self.parse_input(X)
        self._nx = self.X.n_rows
        self._max_core_number = self.X.max_core
        K, xd, _ = self._run_cores(self.X.block, self.X.ids, self.X.rows, self._nx, self._nx)
        self._X_diag = xd
        return K

    def transform(self, X):
        self._method_calling = 3
        check_is_fitted(self, ["X"])
        if X is None:
            raise ValueError("transform input cannot be None")
        Y = self.parse_input(X)
        block = Block.concat(self.X.block, Y.block)
        ids = None if self.X.ids is None else np.concatenate([self.X.ids, Y.ids])
        rows = np.concatenate([self.X.rows, Y.rows + self._nx]).astype(np.int32)
        K, xd, yd = self._run_cores(block, ids, rows, self._nx + Y.n_rows, self._nx)
        self._X_diag, self._Y_diag = xd, yd
        self._is_transformed = True
        return K

    def diagonal(self):
        check_is_fitted(self, ["X"])
        if not hasattr(self, "_X_diag"):
            _, self._X_diag, _ = self._run_cores(self.X.block, self.X.ids, self.X.rows, self._nx, self._nx,
                                                 want_matrix=False)
        if getattr(self, "_is_transformed", False):
            return self._X_diag, self._Y_diag
        return self._X_diag
