#!/usr/bin/env python
"""One steady-state pass of each BASELINE workload inside a profiler range.

Run under `ncu --profile-from-start off ...` (tools/profile.sh): warm-up passes run
outside the range, then exactly one pass of
  config 2  (10 000 graphs, WL-subtree h=5: relabel, feature block, head GEMM + tail)
  config 2d (the same Gram with every shared column dense -- the tensor-bound GEMM)
  config 3  (5 000 graphs, ShortestPath with labels)
  config 5  (2 000 graphs, ShortestPathAttr d=16)      [--spattr]
is captured.  Numbers printed by this script under a profiler are never bench values."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bench import H, N_GRAPHS, pack_workload  # noqa: E402
from grakel_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="wl,dense,sp", help="comma list of wl,dense,sp,spattr,wloa")
    ap.add_argument("--warmup", type=int, default=2)
    args = ap.parse_args()
    what = set(args.what.split(","))
    eng = _lib.get_engine()
    jobs = []
    if "wl" in what or "dense" in what:
        gp, rp, ci, lab = pack_workload(N_GRAPHS)
        n = N_GRAPHS

        def wl():
            eng.pack(gp, rp, ci, lab)
            if "wl" in what:  # the pass as bench.py times it: gk_wl_gram (asynchronous from the second call on)
                eng.wl_gram(H, out=False, dtype=np.float32)
            else:
                eng.wl_features(H)
            if "dense" in what:
                eng.gram(n, out=False, dtype=np.float32, want_diag=False, dense_all=True)
        jobs.append(wl)
    if "sp" in what:
        from grakel_b200.packing import label_ids, pack
        from bench import gen_list as gen  # workload generator
        b = pack(gen(5000, 60, 0, as_adj=True), "sp", want_weights=True)
        ids, _ = label_ids(b.labels, None, sort_new=False)

        def sp():
            eng.pack(b.graph_ptr, b.row_ptr, b.col_idx, ids, b.weights)
            st = eng.sp_features(with_labels=True)
            eng.gram(b.n_graphs, out=False, dtype=np.float32, stats=st, want_diag=False)
        jobs.append(sp)
    if "spattr" in what:
        from grakel_b200.packing import pack
        from bench import gen_list as gen
        b5 = pack(gen(2000, 40, 0, attr=16, as_adj=True), "sp", need_labels=True, attributes=True, want_weights=True)

        def spattr():
            eng.pack(b5.graph_ptr, b5.row_ptr, b5.col_idx, None, b5.weights, b5.attrs)
            st = eng.spattr_features()
            eng.gram(b5.n_graphs, out=False, dtype=np.float64, stats=st, want_diag=False)
        jobs.append(spattr)
    if "wloa" in what:
        from grakel_b200.packing import label_ids, pack
        from bench import gen_list as gen  # workload generator
        bo = pack(gen(10000, 40, 0), "wloa", len_ok=lambda k: k >= 2)
        ido, _ = label_ids(bo.labels, None, sort_new=True)

        def wloa():
            eng.pack(bo.graph_ptr, bo.row_ptr, bo.col_idx, ido)
            st = eng.wl_oa_features(H)
            eng.gram(bo.n_graphs, out=False, dtype=np.float32, stats=st, want_diag=False)
        jobs.append(wloa)
    for j in jobs:
        for _ in range(args.warmup):
            j()
    eng.sync()
    eng.profiler_range(True)
    for j in jobs:
        j()
    eng.sync()
    eng.profiler_range(False)
    print("profiled one pass of:", sorted(what))


if __name__ == "__main__":
    main()
