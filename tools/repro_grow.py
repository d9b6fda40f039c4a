import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bench import pack_workload, H
from grakel_b200 import _lib
eng = _lib.Engine(0)
sizes = [int(x) for x in sys.argv[1:]] or [14142, 20000, 10000, 30000]
for n in sizes:
    gp, rp, ci, lab = pack_workload(n)
    eng.pack(gp, rp, ci, lab)
    try:
        s = eng.wl_features(H)
        eng.gram(n, out=False, dtype=np.float32, stats=s, want_diag=False)
        print(n, "ok", list(s.level_dims)[:6], s.hash_retries, s.ms_features, flush=True)
    except Exception as e:
        print(n, "FAILED", e, flush=True)
