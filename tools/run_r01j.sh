# 2-GPU check of the round's final engine: weak-scaling bench line at N=2 (CTA-pair GEMM on row blocks) and the
# WL-OA expansion with shared-memory aggregated column counts
O=gpurun_out
(timeout 200 python -m pytest tests/test_wloa.py -m gpu -x -q 2>&1 | tail -4) > $O/r01j_wloa_pytest.log 2>&1 &
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --steps 10 --warmup 3 > $O/r01j_bench_n2.json 2> $O/r01j_bench_n2.err
wait
timeout 120 python - > $O/r01j_wloa_timing.json 2> $O/r01j_wloa_timing.err <<'PY'
import json, sys, numpy as np
sys.path.insert(0, ".")
from grakel_b200 import _lib
from grakel_b200.packing import label_ids, pack
from oracle.gk_oracle import gen  # workload generator only
eng = _lib.get_engine()
b = pack(gen(10000, 40, 0), "wloa", len_ok=lambda k: k >= 2)
ids, _ = label_ids(b.labels, None, sort_new=True)
eng.pack(b.graph_ptr, b.row_ptr, b.col_idx, ids)
def step():
    st = eng.wl_oa_features(5)
    eng.gram(b.n_graphs, out=False, dtype=np.float32, stats=st, want_diag=False)
    return st
for _ in range(3): st = step()
eng.event_record(0)
for _ in range(10): st = step()
eng.event_record(1)
ms = eng.event_elapsed(0, 1) / 10
print(json.dumps({"config2_wloa": {"ms_per_step": ms, "pairs_per_s": 1e8 / (ms * 1e-3), "ms_features": st.ms_features,
                                   "ms_columns+panel": st.ms_panel, "ms_gemm": st.ms_gemm, "ms_tail": st.ms_tail}}))
PY
tail -2 $O/r01j_wloa_pytest.log; cat $O/r01j_wloa_timing.json; head -c 700 $O/r01j_bench_n2.json; echo; tail -3 $O/r01j_bench_n2.err
