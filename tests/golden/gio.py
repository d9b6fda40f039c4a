"""Type-preserving JSON (gzip) encoding of graph lists for the golden fixtures.

A fixture must reproduce the *spelling* of the input the reference saw
(adjacency ndarray vs list-of-lists vs scipy sparse vs the five edge-dictionary
forms), because the reference's behaviour depends on it (graph.py:1542-1709).
"""
import gzip
import json

import numpy as np
from scipy.sparse import csr_matrix, issparse


def _enc_key(k):
    if isinstance(k, tuple):
        return {"t": [_enc_key(x) for x in k]}
    if isinstance(k, (np.integer,)):
        return int(k)
    if isinstance(k, (np.floating,)):
        return float(k)
    return k


def _dec_key(k):
    if isinstance(k, dict):
        return tuple(_dec_key(x) for x in k["t"])
    return k


def _enc_val(v):
    if isinstance(v, np.ndarray):
        return {"nd": v.tolist()}
    if isinstance(v, (np.integer,)):
        return int(v)
    if isinstance(v, (np.floating,)):
        return float(v)
    if isinstance(v, tuple):
        return {"t": [_enc_val(x) for x in v]}
    return v


def _dec_val(v):
    if isinstance(v, dict) and "nd" in v:
        return np.asarray(v["nd"], dtype=float)
    if isinstance(v, dict) and "t" in v:
        return tuple(_dec_val(x) for x in v["t"])
    return v


def enc_graph(g):
    if isinstance(g, np.ndarray):
        return {"fmt": "ndarray", "dtype": str(g.dtype), "data": g.tolist()}
    if issparse(g):
        return {"fmt": "sparse", "data": np.asarray(g.todense()).tolist()}
    if type(g) is list and all(isinstance(r, list) for r in g):
        return {"fmt": "lol", "data": g}
    if type(g) is dict:
        vals = list(g.values())
        if all(isinstance(v, list) for v in vals):
            return {"fmt": "dict_list", "data": [[_enc_key(k), [_enc_key(x) for x in v]] for k, v in g.items()]}
        if all(isinstance(v, dict) for v in vals):
            return {"fmt": "dict_dict",
                    "data": [[_enc_key(k), [[_enc_key(a), _enc_val(w)] for a, w in v.items()]] for k, v in g.items()]}
        return {"fmt": "dict_tuple", "data": [[_enc_key(k), _enc_val(w)] for k, w in g.items()]}
    seq = list(g)
    kind = "set" if isinstance(g, (set, frozenset)) else "list"
    return {"fmt": "tuples_" + kind, "data": [[_enc_val(x) for x in t] for t in seq]}


def dec_graph(e):
    f = e["fmt"]
    if f == "ndarray":
        return np.asarray(e["data"], dtype=e["dtype"])
    if f == "sparse":
        return csr_matrix(np.asarray(e["data"], dtype=float))
    if f == "lol":
        return e["data"]
    if f == "dict_list":
        return {_dec_key(k): [_dec_key(x) for x in v] for k, v in e["data"]}
    if f == "dict_dict":
        return {_dec_key(k): {_dec_key(a): _dec_val(w) for a, w in v} for k, v in e["data"]}
    if f == "dict_tuple":
        return {_dec_key(k): _dec_val(w) for k, w in e["data"]}
    if f.startswith("tuples_"):
        seq = [tuple(_dec_val(x) for x in t) for t in e["data"]]
        return set(seq) if f.endswith("set") else seq
    raise ValueError(f)


def enc_dataset(X):
    out = []
    for x in X:
        x = list(x)
        item = {"g": enc_graph(x[0])}
        if len(x) > 1:
            item["l"] = [[_enc_key(k), _enc_val(v)] for k, v in x[1].items()]
        out.append(item)
    return out


def dec_dataset(E):
    out = []
    for item in E:
        el = [dec_graph(item["g"])]
        if "l" in item:
            el.append({_dec_key(k): _dec_val(v) for k, v in item["l"]})
        out.append(el)
    return out


def dump(path, obj):
    with gzip.open(path, "wt", compresslevel=9) as f:
        json.dump(obj, f, separators=(",", ":"))


def load(path):
    with gzip.open(path, "rt") as f:
        return json.load(f)
