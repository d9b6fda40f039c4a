"""CPU oracle for the WL-subtree / Shortest-Path Gram hot path.

TEST INFRASTRUCTURE ONLY.  This module is a CPU restatement of the reference's
(ysig/GraKeL v0.1.11) algorithm for the path named in BASELINE.json.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import it, and only as the checker or the timed
CPU baseline -- the product (``grakel_b200``) never imports anything from
``oracle/`` and fails loudly when its CUDA library is missing.

Parity pinning: every function here is checked against the *real* reference
(imported from /root/reference in the build container) by
``tests/golden/make_golden.py``; the resulting vectors are committed under
``tests/golden/`` and re-checked by ``tests/test_oracle_golden.py``.

The port deliberately keeps the reference's *cost structure* (per-vertex Python
string/tuple keys, per-level sparse feature matrices, one N x N matrix per WL
level summed at the end) so that timing it is a fair stand-in for timing the
reference itself ("kind": "port" in bench.py).

Citations are file:line in /root/reference/grakel/.
"""
from __future__ import annotations

import numbers
from collections import Counter
from collections.abc import Iterable

import numpy as np
from scipy.sparse import csr_matrix, issparse

INF = float("inf")


# --------------------------------------------------------------------------
# input normalisation  (graph.py:147-230, 912-1053, 1542-1709)
# --------------------------------------------------------------------------
def _looks_like_adjacency(g):
    """graph.py:1542-1585 -- 2-D ndarray, scipy sparse, or list of number lists."""
    if isinstance(g, np.ndarray) and g.ndim == 2:
        return True
    if issparse(g):
        return True
    if type(g) is list and all(
        isinstance(r, list) and all(isinstance(x, numbers.Number) for x in r) for r in g
    ):
        return True
    return False


def _edge_dict_from_any(g):
    """Return (vertex set, {u: {v: w}}) or None.  graph.py:1588-1709.

    Five spellings: {(u,v): w}, {u: [v..]}, {u: {v: w}}, iterable of (u,v),
    iterable of (u,v,w).  Vertices that only ever appear as targets get an
    empty out-list (graph.py:1627-1630 and siblings).
    """
    ed = {}

    def put(u, v, w):
        ed.setdefault(u, {})[v] = w

    if type(g) is dict:
        items = list(g.items())
        if all(type(k) is tuple and len(k) == 2 and isinstance(w, numbers.Number) for k, w in items):
            heads, tails = set(), set()
            for (u, v), w in items:
                heads.add(u)
                tails.add(v)
                put(u, v, w)
            for v in tails - heads:
                ed[v] = {}
            return heads | tails, ed
        if all(isinstance(d, list) for d in g.values()):
            heads, tails = set(), set()
            for u, lst in items:
                heads.add(u)
                tails |= set(lst)
                for v in lst:
                    put(u, v, 1.0)
            for v in tails - heads:
                ed[v] = {}
            return heads | tails, ed
        if all(
            isinstance(d, dict) and all(isinstance(w, numbers.Number) for w in d.values())
            for d in g.values()
        ):
            ed = {u: dict(d) for u, d in items}
            heads = set(ed.keys())
            tails = {v for d in ed.values() for v in d}
            for v in tails - heads:
                ed[v] = {}
            return heads | tails, ed
    if isinstance(g, Iterable) and not isinstance(g, (str, bytes, dict)):
        seq = list(g)
        if all(type(t) is tuple and len(t) == 2 for t in seq):
            heads, tails = set(), set()
            for u, v in seq:
                heads.add(u)
                tails.add(v)
                put(u, v, 1.0)
            for v in tails - heads:
                ed[v] = {}
            return heads | tails, ed
        if all(type(t) is tuple and len(t) == 3 for t in seq):
            heads, tails = set(), set()
            for u, v, w in seq:
                heads.add(u)
                tails.add(v)
                put(u, v, w)
            for v in tails - heads:
                ed[v] = {}
            return heads | tails, ed
    return None


class OGraph:
    """Canonical form used by the oracle.

    kind        'adjacency' | 'dictionary'   (what the user passed; decides
                 the "auto" APSP algorithm, graph.py:652-656)
    A           dense float ndarray (adjacency input only)
    verts       sorted vertex symbols (dictionary input: graph.py:902-905;
                 adjacency input: range(n))
    ed          {u: {v: w}} out-neighbour dictionary
    labels      {vertex symbol (dictionary) or index (adjacency): label}
    """

    def __init__(self, g, labels):
        self.labels = labels
        if _looks_like_adjacency(g):
            A = np.asarray(g.todense()) if issparse(g) else np.asarray(g)
            if A.shape[0] != A.shape[1]:
                raise ValueError("input matrix must be squared")
            self.kind = "adjacency"
            self.A = A
            n = A.shape[0]
            self.verts = list(range(n))
            # graph.py:963-965 / 1198-1201: an edge is an entry > 0
            ed = {i: {} for i in range(n)}
            ii, jj = np.where(A > 0)
            for i, j in zip(ii.tolist(), jj.tolist()):
                ed[i][j] = A[i, j]
            self.ed = ed
        else:
            r = _edge_dict_from_any(g)
            if r is None:
                raise ValueError("Unsupported input type.")
            verts, ed = r
            self.kind = "dictionary"
            self.A = None
            self.verts = sorted(verts)
            self.ed = ed

    # -- labels aligned with the APSP matrix index (graph.py:375-399, 689-772)
    def index_labels(self):
        if not self.labels:
            raise ValueError("Graph does not have any labels for vertices.")
        if self.kind == "adjacency":
            return {i: self.labels[i] for i in range(len(self.verts))}
        return {i: self.labels[v] for i, v in enumerate(self.verts)}

    # -- APSP ------------------------------------------------------------
    def floyd_warshall(self):
        """graph.py:1767-1794 (classic k-ordered relaxation, float64)."""
        if self.kind == "adjacency":
            A = self.A
        else:  # graph.py:1026-1038: adjacency built from the edge dictionary
            n = len(self.verts)
            pos = {v: i for i, v in enumerate(self.verts)}
            A = np.zeros((n, n))
            for u, d in self.ed.items():
                for v, w in d.items():
                    A[pos[u], pos[v]] = w
        n = A.shape[0]
        dist = np.array(A, dtype=float, copy=True)
        dist[dist == 0] = INF
        np.fill_diagonal(dist, 0)
        for k in range(n):
            # row form of the reference's i-loop (graph.py:1790-1792): row k and
            # column k are fixed points of step k, so the whole-matrix update is
            # the same arithmetic, one fp64 add + min per entry.
            dist = np.minimum(dist, dist[:, k : k + 1] + dist[k : k + 1, :])
        return dist

    def dijkstra_all(self):
        """graph.py:658-671 + 1712-1764: SSSP from every vertex of the edge dict."""
        import heapq

        n = len(self.verts)
        pos = {v: i for i, v in enumerate(self.verts)}
        S = np.full((n, n), INF)
        for src in self.verts:
            done = {}
            heap = [(0, 0, src)]
            tie = 1
            best = {src: 0}
            while heap:
                d, _, u = heapq.heappop(heap)
                if u in done:
                    continue
                done[u] = d
                for v, w in self.ed[u].items():
                    nd = d + w
                    if v in done:
                        continue
                    if v not in best or nd < best[v]:
                        best[v] = nd
                        heapq.heappush(heap, (nd, tie, v))
                        tie += 1
            for v, d in done.items():
                S[pos[src], pos[v]] = d
        return S

    def shortest_paths(self, algorithm_type="auto"):
        if algorithm_type == "auto":
            algorithm_type = "floyd_warshall" if self.kind == "adjacency" else "dijkstra"
        if algorithm_type == "floyd_warshall":
            return self.floyd_warshall()
        if algorithm_type == "dijkstra":
            return self.dijkstra_all()
        raise ValueError('Unsupported "algorithm_type"')


def _parse(X, min_len=2):
    """Element handling shared by the kernels (weisfeiler_lehman.py:143-194,
    shortest_path.py:432-466): empty elements are skipped, the rest become
    OGraph objects."""
    if not isinstance(X, Iterable):
        raise TypeError("input must be an iterable\n")
    out = []
    for x in X:
        x = list(x)
        if len(x) == 0:
            continue
        lab = x[1] if len(x) > 1 else {}
        out.append(OGraph(x[0], lab))
    if not out:
        raise ValueError("parsed input is empty")
    return out


# --------------------------------------------------------------------------
# Vertex histogram  (vertex_histogram.py:57-219)
# --------------------------------------------------------------------------
class _VH:
    """One fitted base kernel of a WL level: column dictionary + sparse counts."""

    def fit(self, label_dicts):
        self.cols = {}
        self.X = self._features(label_dicts, self.cols)
        return self

    @staticmethod
    def _features(label_dicts, cols):
        r, c, d = [], [], []
        for gi, L in enumerate(label_dicts):
            for lab, cnt in Counter(L.values()).items():
                j = cols.get(lab)
                if j is None:
                    j = len(cols)
                    cols[lab] = j
                r.append(gi)
                c.append(j)
                d.append(cnt)
        return csr_matrix((d, (r, c)), shape=(len(label_dicts), len(cols)), dtype="float64")

    def gram(self):
        return self.X.dot(self.X.T).toarray()  # vertex_histogram.py:177,181-182

    def transform(self, label_dicts):
        cols = dict(self.cols)  # vertex_histogram.py:84 (unseen labels get columns >= D_fit)
        self.Y = self._features(label_dicts, cols)
        return self.Y[:, : self.X.shape[1]].dot(self.X.T).toarray()  # :179

    def xdiag(self):
        return np.asarray(self.X.multiply(self.X).sum(axis=1)).ravel()

    def ydiag(self):
        return np.asarray(self.Y.multiply(self.Y).sum(axis=1)).ravel()


# --------------------------------------------------------------------------
# Edge histogram  (edge_histogram.py:23-212)
# --------------------------------------------------------------------------
class EHOracle:
    """Counts of the VALUES of each element's edge-label dictionary (edge_histogram.py:76-112);
    the graph itself is never consulted.  Elements are [graph, node_labels, edge_labels]."""

    def __init__(self, normalize=False):
        self.normalize = normalize

    @staticmethod
    def _label_dicts(X):
        out = []
        for x in X:
            x = list(x)
            if len(x) == 0:
                continue
            if len(x) != 3:
                raise TypeError("each element of X must be either a graph object or a list with at least a graph like "
                                "object and node labels dict \n")
            out.append(dict(enumerate(x[2].values())))
        if not out:
            raise ValueError("parsed input is empty")
        return out

    def fit_transform(self, X):
        self.vh = _VH().fit(self._label_dicts(X))
        K = self.vh.gram()
        self.xdiag = self.vh.xdiag()
        if self.normalize:
            with np.errstate(divide="ignore", invalid="ignore"):
                K = K / np.sqrt(np.outer(self.xdiag, self.xdiag))
        return K

    def transform(self, Y):
        K = self.vh.transform(self._label_dicts(Y))
        self.ydiag = self.vh.ydiag()
        if self.normalize:
            with np.errstate(divide="ignore", invalid="ignore"):
                K = K / np.sqrt(np.outer(self.ydiag, self.xdiag))
        return K


# --------------------------------------------------------------------------
# Weisfeiler-Lehman subtree  (weisfeiler_lehman.py:117-555)
# --------------------------------------------------------------------------
class WLOracle:
    def __init__(self, n_iter=5, normalize=False, n_jobs=None, base="subtree"):
        if type(n_iter) is not int or n_iter <= 0:
            raise TypeError("'n_iter' must be a positive integer")
        self.h = n_iter
        self.normalize = normalize
        # base kernel fitted on every level's relabelled graphs (weisfeiler_lehman.py:260-270):
        # "subtree" = VertexHistogram (default), "edge_histogram", "shortest_path"
        if base not in ("subtree", "edge_histogram", "shortest_path"):
            raise ValueError("unknown base kernel")
        self.base = base
        # weisfeiler_lehman.py:279-285: with n_jobs the reference hands one task per level
        # (base kernel fit_transform) to a joblib *threading* pool; same structure here.
        self.n_jobs = n_jobs

    def _make_base(self):
        return EHOracle() if self.base == "edge_histogram" else SPOracle(with_labels=True)

    @staticmethod
    def _level_elements(Gs, L, extras):
        """What the reference hands to the base kernel of a level: (edge dictionary, level labels)
        + the edge labels when the element had them (weisfeiler_lehman.py:218, 256)."""
        out = []
        for g, l, e in zip(Gs, L, extras):
            ed = {u: dict(d) for u, d in g.ed.items()}
            for v in l:
                ed.setdefault(v, {})
            out.append([ed, dict(l)] + ([e] if e is not None else []))
        return out

    @staticmethod
    def _signature(own, nbr_labels):
        # weisfeiler_lehman.py:235-239 -- own label, then the sorted multiset of
        # out-neighbour labels.  A tuple keeps exactly the information of the
        # reference's string credential.
        return (own, tuple(sorted(nbr_labels)))

    def fit_transform(self, X, return_levels=False):
        Gs = _parse(X)
        self._fit_graphs = Gs
        L = [dict(g.labels) for g in Gs]
        # level 0: weisfeiler_lehman.py:199-206
        alphabet = sorted({v for l in L for v in l.values()})
        inv0 = {lab: i for i, lab in enumerate(alphabet)}
        self.inv = {0: inv0}
        count = len(inv0)
        L = [{v: inv0[lab] for v, lab in l.items()} for l in L]
        pool = None
        if self.n_jobs is not None and self.n_jobs != 1:
            from concurrent.futures import ThreadPoolExecutor
            import os as _os
            pool = ThreadPoolExecutor(max_workers=self.n_jobs if self.n_jobs > 0 else (_os.cpu_count() or 1))

        extras = []
        for x in X:  # edge labels travel unchanged to every level (weisfeiler_lehman.py:157-169)
            x = list(x)
            if len(x) == 0:
                continue
            extras.append(x[2] if len(x) > 2 else None)
        self._fit_extras = extras

        def level_task(Lc):
            if self.base == "subtree":
                vh = _VH().fit(Lc)
                return vh, vh.gram()
            bk = self._make_base()
            return bk, bk.fit_transform(self._level_elements(Gs, Lc, extras))

        tasks = [pool.submit(level_task, L) if pool else level_task(L)]
        level_labels = [[dict(l) for l in L]] if return_levels else None
        for it in range(1, self.h + 1):  # weisfeiler_lehman.py:223-258
            # The reference collects the credential set, sorts it and numbers it (:243-246);
            # K only depends on the partition, so ids are handed out at first sight here
            # (one pass, no sort) -- the port must not be slower than what it stands in for.
            inv = {}
            newL = []
            for g, l in zip(Gs, L):
                ed = g.ed
                nl = {}
                for v, own in l.items():  # every labelled vertex, sinks included (:230-234)
                    sig = (own, tuple(sorted([l[n] for n in ed.get(v, ())])))
                    i = inv.get(sig)
                    if i is None:
                        i = inv[sig] = count + len(inv)
                    nl[v] = i
                newL.append(nl)
            count += len(inv)
            self.inv[it] = inv
            L = newL
            if return_levels:
                level_labels.append([dict(l) for l in L])
            tasks.append(pool.submit(level_task, L) if pool else level_task(L))
        done = [t.result() if pool else t for t in tasks]
        if pool:
            pool.shutdown()
        self.levels = [vh for vh, _ in done]
        Ks = [k for _, k in done]
        del done, tasks
        K = np.sum(Ks, axis=0)  # weisfeiler_lehman.py:270
        self.level_labels = level_labels
        self.xdiag = np.diagonal(K).copy()
        if self.normalize:  # :324-327
            with np.errstate(divide="ignore", invalid="ignore"):
                K = np.nan_to_num(K / np.sqrt(np.outer(self.xdiag, self.xdiag)))
        if return_levels:
            return K, level_labels
        return K

    def transform(self, Y):
        Gs = _parse(Y)
        L = [dict(g.labels) for g in Gs]
        inv0 = self.inv[0]
        nl = len(inv0)
        fresh = sorted({v for l in L for v in l.values() if v not in inv0})
        new0 = {lab: i for i, lab in enumerate(fresh, nl)}  # weisfeiler_lehman.py:417-418
        L = [{v: (inv0[lab] if lab in inv0 else new0[lab]) for v, lab in l.items()} for l in L]
        yextras = []
        for x in Y:
            x = list(x)
            if len(x) == 0:
                continue
            yextras.append(x[2] if len(x) > 2 else None)

        def level_transform(it, Lc):
            if self.base == "subtree":
                return self.levels[it].transform(Lc)
            return self.levels[it].transform(self._level_elements(Gs, Lc, yextras))

        Ks = [level_transform(0, L)]
        for it in range(1, self.h + 1):  # :435-476
            nl += len(self.inv[it])
            inv = self.inv[it]
            sigs, unseen = [], set()
            for g, l in zip(Gs, L):
                s = {}
                for v in l.keys():
                    s[v] = self._signature(l[v], [l[n] for n in g.ed.get(v, {}).keys()])
                    if s[v] not in inv:
                        unseen.add(s[v])
                sigs.append(s)
            new = {sig: i for i, sig in enumerate(sorted(unseen), nl)}
            L = [{v: (inv[s] if s in inv else new[s]) for v, s in sg.items()} for sg in sigs]
            Ks.append(level_transform(it, L))
        K = np.sum(Ks, axis=0)
        self.ydiag = np.sum([lv.ydiag() if self.base == "subtree" else lv.ydiag for lv in self.levels], axis=0)
        if self.normalize:  # :494-498
            with np.errstate(divide="ignore", invalid="ignore"):
                K = np.nan_to_num(K / np.sqrt(np.outer(self.ydiag, self.xdiag)))
        return K


# --------------------------------------------------------------------------
# Weisfeiler-Lehman optimal assignment  (weisfeiler_lehman_optimal_assignment.py:19-481)
# --------------------------------------------------------------------------
class WLOAOracle:
    """WL-OA: the WL label hierarchy + histogram intersection.

    parse_input (:78-229) relabels like WeisfeilerLehman but numbers the labels of ALL levels in one
    running counter and records each label's parent (the vertex's previous label, :221-229); a graph's
    vector Hs[j] counts, for every vertex, each label on the path from its last-level label up to the
    root (:203-209) -- i.e. the number of vertices of j that carry that label at that label's level.
    K[i, j] = sum_c min(Hs[i, c], Hs[j, c]) (:257-266); transform slices Hs_y[:, :X.shape[1]] (:433) and
    Y_diag uses every Y column (:459-461)."""

    def __init__(self, n_iter=5, normalize=False):
        if type(n_iter) is not int or n_iter <= 0:
            raise TypeError("'n_iter' must be a positive integer")
        self.h = n_iter
        self.normalize = normalize

    def _vectors(self, Gs, L, width):
        """:203-209 / :419-425 -- walk every vertex's label up the hierarchy ('omega' is 1 throughout)."""
        Hs = np.zeros((len(Gs), width))
        for j, l in enumerate(L):
            for lab in l.values():
                cur = lab
                while cur is not None:
                    Hs[j, cur] += 1
                    cur = self.parent[cur]
        return Hs

    @staticmethod
    def _intersect(A, B):
        K = np.empty((A.shape[0], B.shape[0]))
        for i in range(A.shape[0]):
            K[i] = np.minimum(A[i][None, :], B).sum(axis=1)  # :263 / :437
        return K

    def fit_transform(self, X):
        Gs = _parse(X)
        L = [dict(g.labels) for g in Gs]
        self._fit_graphs = Gs
        self.parent = {}
        inv0 = {lab: i for i, lab in enumerate(sorted({v for l in L for v in l.values()}))}  # :157-161
        for i in inv0.values():
            self.parent[i] = None  # children of 'root'
        count = len(inv0)
        self.inv = {0: inv0}
        L = [{v: inv0[lab] for v, lab in l.items()} for l in L]
        for it in range(1, self.h + 1):  # :173-200
            sigs, fresh = [], set()
            for g, l in zip(Gs, L):
                s = {v: WLOracle._signature(l[v], [l[n] for n in nb.keys()]) for v, nb in g.ed.items()}  # keys of the edge dictionary only
                fresh.update(s.values())
                sigs.append(s)
            inv = {}
            for sig in sorted(fresh):  # :186-190 (sorted by credential; the order does not enter K)
                inv[sig] = count
                self.parent[count] = sig[0]
                count += 1
            L = [{v: inv[s] for v, s in sg.items()} for sg in sigs]
            self.inv[it] = inv
        self.n_fit_labels = count
        self.X = self._vectors(Gs, L, count + 1)  # + the (always empty) 'root' column, :204
        K = self._intersect(self.X, self.X)
        self.xdiag = np.diagonal(K).copy()
        if self.normalize:  # :268-272
            with np.errstate(divide="ignore", invalid="ignore"):
                K = np.nan_to_num(K / np.sqrt(np.outer(self.xdiag, self.xdiag)))
        return K

    def transform(self, Y):
        Gs = _parse(Y)
        L = [dict(g.labels) for g in Gs]
        inv0 = self.inv[0]
        parent = dict(self.parent)
        count = self.n_fit_labels  # :355
        new0 = {}
        for lab in sorted({v for l in L for v in l.values() if v not in inv0}):  # :356-359
            new0[lab] = count
            parent[count] = None
            count += 1
        L = [{v: (inv0[lab] if lab in inv0 else new0[lab]) for v, lab in l.items()} for l in L]
        for it in range(1, self.h + 1):  # :370-400
            inv = self.inv[it]
            sigs, unseen = [], set()
            for g, l in zip(Gs, L):
                s = {v: WLOracle._signature(l[v], [l[n] for n in nb.keys()]) for v, nb in g.ed.items()}  # keys of the edge dictionary only
                unseen.update(x for x in s.values() if x not in inv)
                sigs.append(s)
            new = {}
            for sig in sorted(unseen):
                new[sig] = count
                parent[count] = sig[0]
                count += 1
            L = [{v: (inv[s] if s in inv else new[s]) for v, s in sg.items()} for sg in sigs]
        keep, self.parent = self.parent, parent
        Hs = self._vectors(Gs, L, count + 1)
        self.parent = keep
        K = self._intersect(Hs[:, : self.X.shape[1]], self.X)  # :433-437
        self.ydiag = Hs.sum(axis=1)  # :459-461 (min of a row with itself, all columns)
        if self.normalize:  # :440-444
            with np.errstate(divide="ignore", invalid="ignore"):
                K = np.nan_to_num(K / np.sqrt(np.outer(self.ydiag, self.xdiag)))
        return K


def wl_partitions(level_labels):
    """Canonical (first-occurrence) renumbering of each level's labels, vertices
    taken graph by graph in sorted-vertex order.  Two implementations agree on
    the WL partition iff these arrays are equal (SURVEY.md 8c parity rule)."""
    out = []
    for per_graph in level_labels:
        ren, flat = {}, []
        for l in per_graph:
            for v in sorted(l.keys()):
                flat.append(ren.setdefault(l[v], len(ren)))
        out.append(np.asarray(flat, dtype=np.int64))
    return out


# --------------------------------------------------------------------------
# Shortest path (labelled / unlabelled)  (shortest_path.py:167-515)
# --------------------------------------------------------------------------
class SPOracle:
    def __init__(self, with_labels=True, algorithm_type="auto", normalize=False):
        if algorithm_type not in ("auto", "floyd_warshall", "dijkstra"):
            raise ValueError('Unsupported "algorithm_type"')
        self.with_labels = with_labels
        self.algorithm_type = algorithm_type
        self.normalize = normalize

    def _counts(self, X, enum, frozen=None):
        """shortest_path.py:468-490: ordered pairs u != v with finite distance,
        key (l(u), l(v), d) or d; first-seen column numbering."""
        rows = []
        for g in _parse(X, min_len=1):
            S = g.shortest_paths(self.algorithm_type)
            lab = g.index_labels() if self.with_labels else None
            cnt = {}
            n = S.shape[0]
            for u in range(n):
                for v in range(n):
                    if u == v or S[u, v] == INF:
                        continue
                    key = (lab[u], lab[v], S[u, v]) if self.with_labels else S[u, v]
                    if frozen is not None and key in frozen:
                        j = frozen[key]
                    else:
                        j = enum.get(key)
                        if j is None:
                            j = len(enum) + (len(frozen) if frozen is not None else 0)
                            enum[key] = j
                    cnt[j] = cnt.get(j, 0) + 1
            rows.append(cnt)
        return rows

    @staticmethod
    def _dense(rows, D):
        phi = np.zeros((len(rows), D))
        for i, r in enumerate(rows):
            for j, c in r.items():
                phi[i, j] = c
        return phi

    def fit_transform(self, X):
        self.enum = {}
        rows = self._counts(X, self.enum)
        self.phi_x = self._dense(rows, len(self.enum))
        K = np.dot(self.phi_x, self.phi_x.T)  # shortest_path.py:404
        self.xdiag = np.diagonal(K).copy()
        if self.normalize:  # :407-408 (no nan_to_num here)
            with np.errstate(divide="ignore", invalid="ignore"):
                K = K / np.sqrt(np.outer(self.xdiag, self.xdiag))
        return K

    def transform(self, Y):
        yenum = {}
        rows = self._counts(Y, yenum, frozen=self.enum)
        phi_y = self._dense(rows, len(self.enum) + len(yenum))
        K = np.dot(phi_y[:, : len(self.enum)], self.phi_x.T)  # :312
        self.ydiag = np.sum(np.square(phi_y), axis=1)  # :365 (all Y columns)
        if self.normalize:
            with np.errstate(divide="ignore", invalid="ignore"):
                K = K / np.sqrt(np.outer(self.ydiag, self.xdiag))
        return K


# --------------------------------------------------------------------------
# Shortest path on node attributes  (shortest_path.py:16-164)
# --------------------------------------------------------------------------
class SPAttrOracle:
    """k(x,y) = sum_{i!=j} sum_{k!=m} [Sx[i,j] == Sy[k,m] < inf] <a_i,a_k><a_j,a_m>
    (shortest_path.py:151-162, metric = np.dot).  Evaluated per pair with the
    distance-grouped form  sum_d <F_x[d], F_y[d]>,  F[d] = sum_{(i,j):S=d} a_i (x) a_j,
    which is the same bilinear sum re-associated (fp64; SURVEY.md 8a row a20
    measured 1e-15 relative against the 4-deep loop)."""

    def __init__(self, algorithm_type="auto", normalize=False):
        self.algorithm_type = algorithm_type
        self.normalize = normalize

    def _maps(self, X):
        out = []
        for g in _parse(X):
            S = g.shortest_paths(self.algorithm_type)
            lab = g.index_labels()
            n = S.shape[0]
            A = np.asarray([np.asarray(lab[i], dtype=float) for i in range(n)])
            F = {}
            off = ~np.eye(n, dtype=bool)
            for d in np.unique(S[off & np.isfinite(S)]):
                M = ((S == d) & off).astype(float)
                F[float(d)] = A.T @ M @ A
            out.append(F)
        return out

    @staticmethod
    def _pair(Fx, Fy):
        return float(sum(np.vdot(Fx[d], Fy[d]) for d in Fx.keys() & Fy.keys()))

    def pair_bruteforce(self, gx, gy):
        """The reference's literal quadruple loop, for tiny cross-checks."""
        (Sx, ax), (Sy, ay) = gx, gy
        k = 0.0
        for i in range(Sx.shape[0]):
            for j in range(Sx.shape[0]):
                if i == j:
                    continue
                for p in range(Sy.shape[0]):
                    for q in range(Sy.shape[0]):
                        if p == q:
                            continue
                        if Sx[i, j] == Sy[p, q] and Sx[i, j] != INF:
                            k += np.dot(ax[i], ay[p]) * np.dot(ax[j], ay[q])
        return k

    def fit_transform(self, X):
        self.Fx = self._maps(X)
        n = len(self.Fx)
        K = np.zeros((n, n))
        for i in range(n):
            for j in range(i, n):
                K[i, j] = K[j, i] = self._pair(self.Fx[i], self.Fx[j])
        self.xdiag = np.diagonal(K).copy()
        if self.normalize:
            with np.errstate(divide="ignore", invalid="ignore"):
                K = K / np.sqrt(np.outer(self.xdiag, self.xdiag))
        return K

    def transform(self, Y):
        Fy = self._maps(Y)
        K = np.array([[self._pair(fy, fx) for fx in self.Fx] for fy in Fy])
        self.ydiag = np.array([self._pair(f, f) for f in Fy])
        if self.normalize:
            with np.errstate(divide="ignore", invalid="ignore"):
                K = K / np.sqrt(np.outer(self.ydiag, self.xdiag))
        return K


# --------------------------------------------------------------------------
# Seeded synthetic generator (used by every golden and by bench)
# --------------------------------------------------------------------------
def gen(N, nbar, seed, nl=7, attr=0, as_adj=False):
    rs = np.random.RandomState(seed)
    out = []
    for _ in range(N):
        n = int(rs.randint(nbar // 2, nbar + nbar // 2 + 1))
        p = 4.0 / (n - 1)
        iu = np.triu_indices(n, 1)
        m = rs.rand(len(iu[0])) < p
        a, b = iu[0][m], iu[1][m]
        if attr:
            L = {i: rs.rand(attr) for i in range(n)}
        else:
            L = {i: int(rs.randint(nl)) for i in range(n)}
        if as_adj:
            A = np.zeros((n, n))
            A[a, b] = 1.0
            A = A + A.T
            out.append([A, L])
        else:
            g = {}
            for x, y in zip(a.tolist(), b.tolist()):
                g[(x, y)] = 1
                g[(y, x)] = 1
            out.append([g, L])
    return out


def gen_edge_labelled(N, nbar, seed, nl=5, n_el=3):
    """Like gen(), plus an edge-label dictionary {(a, b): l, (b, a): l} per graph (third element)."""
    rs = np.random.RandomState(seed)
    out = []
    for _ in range(N):
        n = int(rs.randint(nbar // 2, nbar + nbar // 2 + 1))
        p = 4.0 / max(n - 1, 1)
        iu = np.triu_indices(n, 1)
        m = rs.rand(len(iu[0])) < p
        g, el = {}, {}
        for a, b in zip(iu[0][m].tolist(), iu[1][m].tolist()):
            l = int(rs.randint(n_el))
            g[(a, b)] = 1
            g[(b, a)] = 1
            el[(a, b)] = l
            el[(b, a)] = l
        if not g:
            g[(0, 1)] = 1
            g[(1, 0)] = 1
            el[(0, 1)] = 0
            el[(1, 0)] = 0
        L = {i: int(rs.randint(nl)) for i in range(n)}
        out.append([g, L, el])
    return out


def oa_sets():
    """Fit / transform sets of the WL-OA goldens (tests/golden/make_golden_oa.py): seeded sparse graphs
    (two labels, ~60 % of the ER edges removed: paths, small trees, isolated vertices -- credentials stay
    shared between graphs at every level), a few one-directional edges, plus level-0 labels the fitted
    set has never seen."""
    X = gen(40, 12, 11, nl=2)
    rs = np.random.RandomState(5)
    for i, (g, _l) in enumerate(X):
        for (a, b) in [e for e in sorted(g) if e[0] < e[1]]:
            r = rs.rand()
            if r < 0.6:
                del g[(a, b)], g[(b, a)]
            elif r < 0.65 and i % 3 == 0:
                del g[(b, a)]  # directed
    fit, new = X[:28], X[28:]
    for _g, l in new[::2]:
        for v in list(l)[::4]:
            l[v] = 7 + int(rs.randint(2))  # labels 7, 8 do not occur in `fit`
    return fit, new


# --------------------------------------------------------------------------
# TU-format dataset files  (datasets/base.py:135-290)
# --------------------------------------------------------------------------
def read_data_oracle(path, name, is_symmetric=False, prefer_attr_nodes=False, produce_labels_nodes=False):
    """Restatement of `read_data`: -> ([[set of (u, v), {node: label}, {(u, v): edge label}], ...], classes or
    None).  Graph ids and node ids are the 1-based ids of the files; an edge line belongs to the graph of its
    source (:213-215), `is_symmetric` adds the reverse to the graph of the target (:216-218)."""
    import os as _os
    if _os.path.isdir(_os.path.join(path, name)):  # ./<name>/<name>_A.txt (:181-191)
        path = _os.path.join(path, name)
    base = _os.path.join(path, name + "_")

    def lines(suffix):
        fn = base + suffix
        if not _os.path.exists(fn):
            return None
        with open(fn) as f:
            return [ln[:-1] if ln.endswith("\n") else ln for ln in f]

    ngc, graphs, nlab, elab = {}, {}, {}, {}
    for i, ln in enumerate(lines("graph_indicator.txt"), 1):  # :203-211
        g = int(ln)
        ngc[i] = g
        graphs.setdefault(g, set()); nlab.setdefault(g, {}); elab.setdefault(g, {})
    elc = {}
    for i, ln in enumerate(lines("A.txt"), 1):  # :214-220
        a, b = (int(x) for x in ln.replace(" ", "").split(","))
        elc[i] = (a, b)
        graphs[ngc[a]].add((a, b))
        if is_symmetric:
            graphs[ngc[b]].add((b, a))
    attr, nl = lines("node_attributes.txt"), lines("node_labels.txt")
    if prefer_attr_nodes and attr is not None:  # :223-232
        for i, ln in enumerate(attr, 1):
            nlab[ngc[i]][i] = [float(x) for x in ln.replace(" ", "").split(",")]
    elif nl is not None:  # :234-240
        for i, ln in enumerate(nl, 1):
            nlab[ngc[i]][i] = int(ln)
    elif produce_labels_nodes:  # :241-243
        for g in graphs:
            nlab[g] = dict(Counter(s for (s, d) in graphs[g] if s != d))
    el = lines("edge_labels.txt")
    if el is not None:  # :259-267
        for i, ln in enumerate(el, 1):
            a, b = elc[i]
            elab[ngc[a]][(a, b)] = int(ln)
            if is_symmetric:
                elab[ngc[b]][(b, a)] = int(ln)
    data = [[graphs[g], nlab[g], elab[g]] for g in range(1, len(graphs) + 1)]  # :274-276
    cl = lines("graph_labels.txt")
    return data, (None if cl is None else np.array([int(x) for x in cl], dtype=int))


def tu_digest(elements, mode):
    """Canonical digest of what a kernel sees in read_data's elements: per graph the sorted (node, label) pairs of
    its vertex set and its sorted edge list.  mode 'wl': vertices = label keys; 'sp': vertices = edge endpoints."""
    import hashlib
    h = hashlib.sha1()
    for g, l, _e in elements:
        verts = sorted(l) if mode == "wl" else sorted({x for e in g for x in e})
        edges = sorted((a, b) for (a, b) in g if mode != "wl" or a in l)
        lab = [l[v] for v in verts]
        lab = [int(x) if not isinstance(x, list) else tuple(float(y) for y in x) for x in lab]
        h.update(repr((verts, lab, edges)).encode())
    return h.hexdigest()
