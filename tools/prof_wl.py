import sys, os
sys.path.insert(0, '.')
import numpy as np
from bench import pack_workload, H, N_GRAPHS
from grakel_b200 import _lib
eng = _lib.get_engine()
gp, rp, ci, lab = pack_workload(N_GRAPHS)
eng.pack(gp, rp, ci, lab)
for i in range(3):
    st = eng.wl_features(H)
os.environ["GRAKEL_B200_PROF"] = "1"
st = eng.wl_features(H)
print("ms_features", st.ms_features)
