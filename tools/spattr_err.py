"""SP-attr Gram accuracy of the 3xTF32 tensor-core path against the fp64 CUDA-core Gram, by k-chunk length."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
code = r'''
import os, sys, numpy as np
sys.path.insert(0, %r)
from bench import gen_list
from grakel_b200 import ShortestPathAttr
X = gen_list(200, 40, 0, attr=16, as_adj=True)
K = ShortestPathAttr().fit_transform(X)
os.environ["GRAKEL_B200_SPATTR_F64"] = "1"
K64 = ShortestPathAttr().fit_transform(X)
rel = (K - K64) / np.abs(K64)
print("chunk", os.environ.get("GRAKEL_B200_SPATTR_CHUNK"), "max rel", float(np.max(np.abs(rel))), "mean rel (signed)", float(rel.mean()), "sym", bool(np.array_equal(K, K.T)))
''' % ROOT
for chunk in ("1", "2", "4", "8", "24", "1000"):
    env = dict(os.environ, GRAKEL_B200_SPATTR_CHUNK=chunk)
    env.pop("GRAKEL_B200_SPATTR_F64", None)
    subprocess.run([sys.executable, "-c", code], env=env)
