"""Generate the golden fixtures by running the REAL reference (ysig/GraKeL).

Runs only in the build container, where /root/reference exists.  The reference
needs its Cython helpers built to be importable, so point GRAKEL_REF at a
writable, built copy:

    cp -r /root/reference /tmp/ref && chmod -R u+w /tmp/ref
    (cd /tmp/ref && python setup.py build_ext --inplace)
    GRAKEL_REF=/tmp/ref python tests/golden/make_golden.py            # small fixtures
    GRAKEL_REF=/tmp/ref python tests/golden/make_golden.py --big      # config 2/3 checksums (~15 min, 12 GB)

Nothing here is imported by the product or by the GPU tests; the fixtures it
writes (tests/golden/*.json.gz, *.npz) are what travels.
"""
import hashlib
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.environ.get("GRAKEL_REF", "/tmp/ref"))

import gio  # noqa: E402
from scipy.sparse import csr_matrix  # noqa: E402

import grakel  # noqa: E402  (the reference)
from grakel import ShortestPath, ShortestPathAttr, WeisfeilerLehman  # noqa: E402
from grakel.datasets import generate_dataset  # noqa: E402
from grakel.datasets.base import read_data  # noqa: E402

from oracle.gk_oracle import gen  # noqa: E402  (the seeded generator of SURVEY 8d)


def summ(K):
    K = np.ascontiguousarray(K, dtype=np.float64)
    return {
        "shape": list(K.shape),
        "sum": float(K.sum()),
        "trace": float(np.trace(K)) if K.shape[0] == K.shape[1] else None,
        "max": float(K.max()),
        "sha1": hashlib.sha1(K.tobytes()).hexdigest(),
    }


def mat(K):
    """Exact text form of a float64 matrix (repr round-trips)."""
    return np.asarray(K, dtype=np.float64).tolist()


# -- A. input spellings ---------------------------------------------------
def spelling_cases():
    A = np.array([[0, 1, 0, 0, 2], [1, 0, 1, 0, 0], [0, 1, 0, 3, 0], [0, 0, 3, 0, 1], [2, 0, 0, 1, 0]], dtype=float)
    Adir = np.array([[0, 1, 0, 0], [0, 0, 2, 0], [0, 0, 0, 1], [0, 0, 0, 0]], dtype=float)  # chain, vertex 3 is a sink
    L5 = {0: "a", 1: "b", 2: "a", 3: "c", 4: "b"}
    L4 = {0: "a", 1: "b", 2: "a", 3: "b"}
    cases = {}
    cases["ndarray_sym"] = [[A, L5], [Adir, L4], [A[:3, :3].copy(), {0: "a", 1: "a", 2: "b"}]]
    cases["lol"] = [[A.tolist(), L5], [Adir.tolist(), L4]]
    cases["sparse"] = [[csr_matrix(A), L5], [csr_matrix(Adir), L4]]
    cases["int_ndarray"] = [[A.astype(int), L5], [Adir.astype(int), L4]]
    cases["dict_tuple"] = [
        [{("x", "y"): 1.0, ("y", "x"): 1.0, ("y", "z"): 2.0, ("z", "y"): 2.0}, {"x": 0, "y": 1, "z": 0}],
        [{(1, 2): 1, (2, 3): 1, (3, 1): 1, (3, 4): 1}, {1: 0, 2: 0, 3: 1, 4: 1}],  # directed, 4 is a sink
    ]
    cases["dict_list"] = [
        [{"p": ["q", "r"], "q": ["p"], "r": ["p", "q"]}, {"p": 5, "q": 6, "r": 5}],
        [{0: [1], 1: [0, 2], 2: [1, 3]}, {0: 1, 1: 1, 2: 2, 3: 2}],  # 3 only as a target
    ]
    cases["dict_dict"] = [
        [{0: {1: 1.0, 2: 2.0}, 1: {0: 1.0}, 2: {0: 2.0, 3: 1.0}, 3: {2: 1.0}}, {0: "u", 1: "v", 2: "u", 3: "v"}],
        [{10: {20: 1}, 20: {10: 1, 30: 1}, 30: {20: 1}}, {10: "u", 20: "u", 30: "v"}],
    ]
    cases["tuples2"] = [
        [[(0, 1), (1, 0), (1, 2), (2, 1), (2, 0), (0, 2)], {0: 1, 1: 2, 2: 3}],
        [[(5, 6), (6, 5), (6, 7)], {5: 1, 6: 1, 7: 2}],
    ]
    cases["tuples3"] = [
        [[(0, 1, 2.0), (1, 0, 2.0), (1, 2, 1.0), (2, 1, 1.0)], {0: 1, 1: 2, 2: 1}],
        [[("a", "b", 1), ("b", "c", 1), ("c", "a", 4)], {"a": 0, "b": 0, "c": 1}],
    ]
    # mixed spellings + an isolated, labelled vertex (kept by WL, featureless for SP)
    Aiso = np.zeros((4, 4))
    Aiso[0, 1] = Aiso[1, 0] = Aiso[1, 2] = Aiso[2, 1] = 1
    cases["mixed_isolated"] = [[Aiso, {0: 0, 1: 1, 2: 0, 3: 1}], cases["dict_tuple"][1], cases["tuples2"][0],
                               [A, {0: 3, 1: 0, 2: 3, 3: 1, 4: 0}]]
    return cases


def run_all_kernels(X, Y=None, wl_iters=(1, 3)):
    out = {}
    for h in wl_iters:
        for norm in (False, True):
            wl = WeisfeilerLehman(n_iter=h, normalize=norm)
            key = f"wl_h{h}_{'n' if norm else 'u'}"
            out[key] = {"fit_transform": mat(wl.fit_transform(X))}
            if not norm:
                out[key]["D"] = [int(wl.X[i].X.shape[1]) for i in range(h + 1)]
            if Y is not None:
                out[key]["transform"] = mat(wl.transform(Y))
    for wlab in (True, False):
        for alg in ("auto", "floyd_warshall", "dijkstra"):
            for norm in (False, True):
                key = f"sp_{'l' if wlab else 'x'}_{alg}_{'n' if norm else 'u'}"
                try:
                    sp = ShortestPath(with_labels=wlab, algorithm_type=alg, normalize=norm)
                    with np.errstate(all="ignore"):
                        rec = {"fit_transform": mat(sp.fit_transform(X))}
                        if Y is not None:
                            rec["transform"] = mat(sp.transform(Y))
                    out[key] = rec
                except Exception as e:  # the reference itself fails on this spelling
                    out[key] = {"error": type(e).__name__}
    return out


def main_small():
    import warnings

    warnings.simplefilter("ignore")
    G = {"reference_version": grakel.__version__, "cases": {}}
    for name, X in spelling_cases().items():
        G["cases"][name] = {"X": gio.enc_dataset(X), "out": run_all_kernels(X)}
    gio.dump(os.path.join(HERE, "spellings.json.gz"), G)
    print("spellings:", len(G["cases"]))

    # -- B. bundled MUTAG (the config-1 real-data twin) ----------------------
    cwd = os.getcwd()
    os.chdir(os.path.join(os.environ.get("GRAKEL_REF", "/tmp/ref"), "grakel", "tests", "data"))
    mutag = read_data("MUTAG", with_classes=False).data  # [set of (u,v), node labels, edge labels]
    os.chdir(cwd)
    # re-spell as a sorted list of (u,v) tuples + node labels so the fixture is self-contained
    Xm = [[sorted((int(a), int(b)) for a, b in g[0]), {int(k): int(v) for k, v in g[1].items()}] for g in mutag]
    ks = {}
    for h in (3, 5):
        t = time.perf_counter()
        K = WeisfeilerLehman(n_iter=h).fit_transform(mutag)
        dt = time.perf_counter() - t
        assert np.array_equal(K, WeisfeilerLehman(n_iter=h).fit_transform(Xm))
        ks[f"wl_h{h}"] = K.astype(np.int64)
        print(f"MUTAG WL h={h}", summ(K), f"{dt:.3f}s")
    K = ShortestPath().fit_transform(mutag)
    assert np.array_equal(K, ShortestPath().fit_transform(Xm))
    ks["sp"] = K.astype(np.int64)
    print("MUTAG SP", summ(K))
    idx = np.random.RandomState(42).permutation(len(Xm))
    tr, te = idx[:150].tolist(), idx[150:].tolist()
    wl = WeisfeilerLehman(n_iter=3, normalize=True)
    ks["wl_h3_norm_train"] = wl.fit_transform([Xm[i] for i in tr])
    ks["wl_h3_norm_test"] = wl.transform([Xm[i] for i in te])
    sp = ShortestPath(normalize=True)
    ks["sp_norm_train"] = sp.fit_transform([Xm[i] for i in tr])
    ks["sp_norm_test"] = sp.transform([Xm[i] for i in te])
    ks["split_train"] = np.asarray(tr)
    ks["split_test"] = np.asarray(te)
    np.savez_compressed(os.path.join(HERE, "mutag_out.npz"), **ks)
    gio.dump(os.path.join(HERE, "mutag_graphs.json.gz"), gio.enc_dataset(Xm))

    # -- C. config 1: 188 synthetic graphs, WL h=3 ---------------------------
    X1 = gen(188, 18, 0)
    wl = WeisfeilerLehman(n_iter=3)
    t = time.perf_counter()
    K1 = wl.fit_transform(X1)
    dt = time.perf_counter() - t
    print("config1", summ(K1), [int(wl.X[i].X.shape[1]) for i in range(4)], f"{dt:.3f}s")
    np.savez_compressed(os.path.join(HERE, "config1_out.npz"), K=K1.astype(np.int64),
                        D=np.asarray([int(wl.X[i].X.shape[1]) for i in range(4)]),
                        Knorm=WeisfeilerLehman(n_iter=3, normalize=True).fit_transform(X1))

    # -- D. fit/transform with unseen labels, real & unit weights ------------
    D = {}
    for tag, rw, seed in (("unit", (1, 1), 42), ("intw", (1, 4), 7), ("realw", (0.5, 2.5), 11)):
        tr, te = generate_dataset(n_graphs=40, r_vertices=(4, 18), r_connectivity=(0.2, 0.7), r_weight_edges=rw,
                                  n_graphs_test=12, random_state=seed, features=("nl", 3))
        if tag == "intw":  # integer weights on the same structure
            tr = [(np.rint(a), l) for a, l in tr]
            te = [(np.rint(a), l) for a, l in te]
        tr, te = [list(x) for x in tr], [list(x) for x in te]
        D[tag] = {"X": gio.enc_dataset(tr), "Y": gio.enc_dataset(te), "out": run_all_kernels(tr, te, wl_iters=(2, 4))}
    gio.dump(os.path.join(HERE, "fit_transform.json.gz"), D)

    # -- E. ShortestPathAttr on tiny graphs (the reference's 4-deep loop) -----
    Xa = gen(7, 7, 3, attr=4, as_adj=True)
    Xa_tr, Xa_te = Xa[:5], Xa[5:]
    spa = ShortestPathAttr()
    Ka = spa.fit_transform(Xa_tr)
    Kt = spa.transform(Xa_te)
    spn = ShortestPathAttr(normalize=True)
    gio.dump(os.path.join(HERE, "spattr.json.gz"),
             {"X": gio.enc_dataset(Xa_tr), "Y": gio.enc_dataset(Xa_te), "K": mat(Ka), "Kt": mat(Kt),
              "Kn": mat(spn.fit_transform(Xa_tr)), "Ktn": mat(spn.transform(Xa_te))})
    print("spattr", Ka[0])

    # -- F. config 3 scaled down: 40 graphs nbar=60, adjacency => FW ---------
    X3 = gen(40, 60, 0, as_adj=True)
    sp = ShortestPath()
    K3 = sp.fit_transform(X3)
    np.savez_compressed(os.path.join(HERE, "config3_small_out.npz"), K=K3.astype(np.int64), D=len(sp._enum))
    print("config3_small", summ(K3), len(sp._enum))


def main_big():
    """Full-size checksums + sampled rows for configs 2 and 3 (reference CPU)."""
    rows = [0, 1, 4999, 5000, 9999]
    out = {}
    X3 = gen(5000, 60, 0, as_adj=True)
    t = time.perf_counter()
    sp = ShortestPath()
    K = sp.fit_transform(X3)
    dt = time.perf_counter() - t
    out["config3"] = dict(summ(K), seconds=dt, D=len(sp._enum), cpu_count=os.cpu_count())
    r3 = [r for r in rows if r < 5000]
    np.savez_compressed(os.path.join(HERE, "config3_rows.npz"), rows=np.asarray(r3), K_rows=K[r3].astype(np.int64),
                        diag=np.diagonal(K).astype(np.int64))
    print("config3", out["config3"], flush=True)
    del K, sp, X3
    X2 = gen(10000, 40, 0)
    t = time.perf_counter()
    wl = WeisfeilerLehman(n_iter=5)
    K = wl.fit_transform(X2)
    dt = time.perf_counter() - t
    out["config2"] = dict(summ(K), seconds=dt, D=[int(wl.X[i].X.shape[1]) for i in range(6)],
                          cpu_count=os.cpu_count())
    np.savez_compressed(os.path.join(HERE, "config2_rows.npz"), rows=np.asarray(rows), K_rows=K[rows].astype(np.int64),
                        diag=np.diagonal(K).astype(np.int64))
    print("config2", out["config2"], flush=True)
    gio.dump(os.path.join(HERE, "big_summaries.json.gz"), out)


if __name__ == "__main__":
    if "--big" in sys.argv:
        main_big()
    else:
        main_small()
