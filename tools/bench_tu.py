#!/usr/bin/env python
"""Secondary measurement: TU-format files on disk -> K on the host.

Writes BASELINE config 2 (10 000 graphs) as TU text files into a scratch directory, then times
  host_reference_style : the oracle's restatement of read_data (per-line Python, datasets/base.py:135-290)
                         + the Python packer every fit() of the list path runs   [CPU leg, "port"]
  native_reader        : gk_tu_open / gk_tu_pack / gk_tu_fill through grakel_b200.datasets.read_tu
  files_to_K           : read_tu + WeisfeilerLehman(n_iter=5).fit_transform(block) (float64 K on the host); needs a GPU
"""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import tu_io  # noqa: E402
from grakel_b200.datasets import read_tu  # noqa: E402
from grakel_b200.packing import pack  # noqa: E402
from oracle.gk_oracle import gen, read_data_oracle  # noqa: E402  (workload generator + CPU leg)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    gpu = "--no-gpu" not in sys.argv
    X = tu_io.renumber(gen(n, 40, 0))
    out = {"graphs": n}
    with tempfile.TemporaryDirectory() as d:
        tu_io.write_tu(d, "CFG2", X, classes=[i % 2 for i in range(n)])
        out["file_bytes"] = sum(os.path.getsize(os.path.join(d, f)) for f in os.listdir(d))
        t0 = time.perf_counter()
        data, _ = read_data_oracle(d, "CFG2")
        t1 = time.perf_counter()
        pack(data, "wl", len_ok=lambda k: k >= 2)
        t2 = time.perf_counter()
        out["host_reference_style"] = {"read_data_s": t1 - t0, "pack_s": t2 - t1, "kind": "port, 1 thread"}
        ts = []
        for _ in range(5):
            t = time.perf_counter()
            b = read_tu(d, "CFG2", kernel="WL")
            ts.append(time.perf_counter() - t)
        out["native_reader"] = {"seconds_best": min(ts), "seconds_median": float(np.median(ts)),
                                "vertices": int(b.data.n_vertices), "edges": int(len(b.data.col_idx)),
                                "MB_per_s": out["file_bytes"] / min(ts) / 1e6}
        out["reader_speedup"] = (t2 - t0) / min(ts)
        if gpu:
            from grakel_b200 import WeisfeilerLehman
            est = WeisfeilerLehman(n_iter=5)
            est.fit_transform(read_tu(d, "CFG2", kernel="WL").data)  # warm-up (allocations, pinned buffers)
            ts = []
            for _ in range(3):
                t = time.perf_counter()
                K = est.fit_transform(read_tu(d, "CFG2", kernel="WL").data)
                ts.append(time.perf_counter() - t)
            out["files_to_K"] = {"seconds_best": min(ts), "pairs_per_s": n * n / min(ts), "K_sum": float(K.sum()),
                                 "stats": {k: v for k, v in est.stats_.as_dict().items() if k.startswith("ms_")}}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
