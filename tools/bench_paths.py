#!/usr/bin/env python
"""Secondary measurements (not the driver's bench line): BASELINE configs 3 (ShortestPath,
5 000 graphs, avg 60 nodes) and 5 (ShortestPathAttr, 2 000 graphs, d=16), and WL-OA on the graphs of
config 2, on one GPU,
CSR resident in HBM -> K resident in HBM, CUDA events on the engine's stream."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from grakel_b200 import _lib  # noqa: E402
from grakel_b200.packing import label_ids, pack  # noqa: E402
from oracle.gk_oracle import gen  # noqa: E402  (workload generator only)


def timed(eng, fn, steps=10, warmup=3):
    for _ in range(warmup):
        st = fn()
    eng.event_record(0)
    for _ in range(steps):
        st = fn()
    eng.event_record(1)
    return eng.event_elapsed(0, 1) / steps, st


def main():
    eng = _lib.get_engine()
    out = {}
    # ---- config 3
    X = gen(5000, 60, 0, as_adj=True)
    t = time.perf_counter()
    b = pack(X, "sp", want_weights=True)
    ids, _ = label_ids(b.labels, None, sort_new=False)
    t_pack = time.perf_counter() - t
    eng.pack(b.graph_ptr, b.row_ptr, b.col_idx, ids, b.weights)
    n = b.n_graphs
    sizes = np.diff(b.graph_ptr).astype(np.int64)

    def step():
        st = eng.sp_features(with_labels=True)
        eng.gram(n, out=False, dtype=np.float32, stats=st, want_diag=False)
        return st

    ms, st = timed(eng, step)
    relax = float((sizes ** 3).sum())
    pairs = float((sizes * (sizes - 1)).sum())
    out["config3_sp"] = {
        "graphs": n, "vertices": int(b.graph_ptr[-1]), "edges": int(b.row_ptr[-1]), "ms_per_step": ms,
        "pairs_per_s": n * n / (ms * 1e-3), "ms_features(APSP+histogram)": st.ms_features, "ms_columns+panel": st.ms_panel,
        "ms_gemm": st.ms_gemm, "ms_tail": st.ms_tail, "features_D": int(st.n_columns), "head_columns": int(st.n_dense_columns),
        "threshold_T": int(st.threshold), "fw_minplus_per_s": relax / (st.ms_features * 1e-3),
        "vertex_pairs_per_s": pairs / (st.ms_features * 1e-3),
        "apsp_compulsory_GBps": (4.0 * (b.graph_ptr[-1] + b.row_ptr[-1]) + 4.0 * b.graph_ptr[-1]) / (st.ms_features * 1e-3) / 1e9,
        "host_pack_s": t_pack, "gram_path": int(st.gram_path), "max_count": int(st.max_count)}
    # ---- config 5
    X = gen(2000, 40, 0, attr=16, as_adj=True)
    b = pack(X, "sp", need_labels=True, attributes=True, want_weights=True)
    eng.pack(b.graph_ptr, b.row_ptr, b.col_idx, None, b.weights, b.attrs)
    n = b.n_graphs

    def step5():
        st = eng.spattr_features()
        eng.gram(n, out=False, dtype=np.float64, stats=st, want_diag=False)
        return st

    ms, st = timed(eng, step5, steps=5, warmup=2)
    D = int(st.n_columns)
    out["config5_spattr"] = {"graphs": n, "feature_dim": D, "distance_blocks": int(st.level_dims[0]), "ms_per_step": ms,
                             "pairs_per_s": n * n / (ms * 1e-3), "ms_features": st.ms_features, "ms_gram_fp64": st.ms_gemm,
                             "gram_fp64_tflops": 2.0 * n * n * D / (st.ms_gemm * 1e-3) / 1e12}
    # ---- WL-OA on the graphs of config 2: unary-expanded WL block, same Gram
    X = gen(10000, 40, 0)
    b = pack(X, "wloa", len_ok=lambda k: k >= 2)
    ids, _ = label_ids(b.labels, None, sort_new=True)
    eng.pack(b.graph_ptr, b.row_ptr, b.col_idx, ids)
    n = b.n_graphs

    def step_oa():
        st = eng.wl_oa_features(5)
        eng.gram(n, out=False, dtype=np.float32, stats=st, want_diag=False)
        return st

    ms, st = timed(eng, step_oa)
    from oracle.gk_oracle import WLOAOracle  # CPU leg of this secondary measurement
    m = 150
    t = time.perf_counter()
    WLOAOracle(n_iter=5).fit_transform(X[:m])
    t_cpu = time.perf_counter() - t
    out["config2_wloa"] = {
        "graphs": n, "vertices": int(b.graph_ptr[-1]), "ms_per_step": ms, "pairs_per_s": n * n / (ms * 1e-3),
        "ms_features(WL + unary expansion)": st.ms_features, "ms_columns+panel": st.ms_panel, "ms_gemm": st.ms_gemm,
        "ms_tail": st.ms_tail, "unary_columns": int(st.n_columns), "unary_entries": int(st.n_entries),
        "head_columns": int(st.n_dense_columns), "threshold_T": int(st.threshold), "tail_updates": int(st.tail_updates),
        "cpu_port": {"graphs": m, "seconds": t_cpu, "pairs_per_s": m * m / t_cpu, "kind": "port, 1 thread"}}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
