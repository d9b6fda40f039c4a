"""Per-level profile of wl_fused2 (GRAKEL_B200_PROF) at a given number of graphs: python tools/prof_wl_n.py 28284"""
import os
import sys
sys.path.insert(0, '.')
from bench import pack_workload, H
from grakel_b200 import _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
eng = _lib.get_engine()
eng.pack(*pack_workload(n))
for i in range(3):
    st = eng.wl_features(H)
print("n", n, "ms_features", st.ms_features, "level dims", [st.level_dims[i] for i in range(H + 1)])
os.environ["GRAKEL_B200_PROF"] = "1"
st = eng.wl_features(H)
