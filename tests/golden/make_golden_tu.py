"""Digests of what the REAL reference's read_data returns for its own bundled TU datasets (MUTAG, Cuneiform),
in the canonical form of oracle.gk_oracle.tu_digest -> tests/golden/tu_digest.json.

    GRAKEL_REF=/tmp/grakel_ref python tests/golden/make_golden_tu.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
REF = os.environ.get("GRAKEL_REF", "/tmp/ref")
sys.path.insert(0, REF)

from grakel.datasets.base import read_data  # noqa: E402  (the reference)

from oracle.gk_oracle import tu_digest  # noqa: E402

out = {}
cwd = os.getcwd()
os.chdir(os.path.join(REF, "grakel", "tests", "data"))
for name in ("MUTAG", "Cuneiform"):
    for sym in (False, True):
        for attr in (False, True):
            key = f"{name}_sym{int(sym)}_attr{int(attr)}"
            try:
                d = read_data(name, with_classes=True, is_symmetric=sym, prefer_attr_nodes=attr)
            except ValueError as e:  # Cuneiform's node_labels.txt has two columns: int() fails in the reference too
                out[key] = {"error": "ValueError"}
                continue
            out[key] = {"wl": tu_digest(d.data, "wl"), "sp": tu_digest(d.data, "sp"), "graphs": len(d.data),
                        "classes_sum": int(d.target.sum()),
                        "edge_label_entries": int(sum(len(e[2]) for e in d.data)),
                        "edge_label_sum": int(sum(sum(v for v in e[2].values() if isinstance(v, int)) for e in d.data))}
os.chdir(cwd)
json.dump(out, open(os.path.join(HERE, "tu_digest.json"), "w"), indent=1, sort_keys=True)
print(json.dumps(out, indent=1, sort_keys=True))
