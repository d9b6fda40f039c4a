/* _fastpack -- CPython-level packer for the commonest input spelling of the hot path.
 *
 * The reference turns every element [graph, node labels] of a `fit` into a `Graph` object in pure Python
 * (graph.py:147-230, 982-1053: ~1 ms per graph); grakel_b200.packing.pack does the same normalisation with numpy
 * (~0.1 ms per graph) but still walks each edge in byte code or through generic C iterators.  This module walks
 * the dictionaries with the CPython API directly for the spelling  {(u, v): w}  with integer vertex symbols and
 * dictionary labels, and builds the CSR block of ALL graphs in one call.  Anything else -- other spellings, other
 * symbol types, unlabelled or foreign vertices, empty graphs -- makes it return None and the Python packer takes
 * over, so semantics and error behaviour stay in one place (packing.py).
 *
 * pack_edge_dicts(elements, mode, need_labels) -> None | (graph_ptr, row_ptr, col_idx, weights, labels, any_weight)
 *   elements     list of (graph dict, label dict) pairs
 *   mode         0: vertex set = label keys, a contiguous integer range in key order  (packing mode 'wl')
 *                1: vertex set = the symbols that occur in an edge, sorted            (packing modes 'sp', 'wloa')
 *   graph_ptr / row_ptr / col_idx : bytes of int32 arrays; weights : bytes of float64; labels : list (or None)
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  void* p;
  size_t n, cap, esz;
} vec;

static int vec_push(vec* v, const void* x) {
  if (v->n == v->cap) {
    size_t nc = v->cap ? v->cap * 2 : 1024;
    void* q = realloc(v->p, nc * v->esz);
    if (!q) return -1;
    v->p = q;
    v->cap = nc;
  }
  memcpy((char*)v->p + v->n * v->esz, x, v->esz);
  v->n += 1;
  return 0;
}

static int cmp_i64(const void* a, const void* b) {
  const int64_t x = *(const int64_t*)a, y = *(const int64_t*)b;
  return x < y ? -1 : x > y;
}

/* exact int -> int64; returns 0 when the object is not an exact int or does not fit */
static int as_i64(PyObject* o, int64_t* out) {
  if (!PyLong_CheckExact(o)) return 0;
  int overflow = 0;
  const long long v = PyLong_AsLongLongAndOverflow(o, &overflow);
  if (overflow || (v == -1 && PyErr_Occurred())) { PyErr_Clear(); return 0; }
  *out = (int64_t)v;
  return 1;
}

static int as_weight(PyObject* o, double* out) {
  if (PyFloat_CheckExact(o)) { *out = PyFloat_AS_DOUBLE(o); return 1; }
  if (PyLong_CheckExact(o) || PyBool_Check(o)) {
    const double v = PyLong_AsDouble(o);
    if (v == -1.0 && PyErr_Occurred()) { PyErr_Clear(); return 0; }
    *out = v;
    return 1;
  }
  return 0;
}

static PyObject* pack_edge_dicts(PyObject* self, PyObject* args) {
  PyObject* seq_in;
  int mode, need_labels;
  if (!PyArg_ParseTuple(args, "Oii", &seq_in, &mode, &need_labels)) return NULL;
  PyObject* seq = PySequence_Fast(seq_in, "elements must be a sequence");
  if (!seq) return NULL;
  const Py_ssize_t N = PySequence_Fast_GET_SIZE(seq);
  vec gptr = {0, 0, 0, 4}, src = {0, 0, 0, 4}, dst = {0, 0, 0, 4}, wts = {0, 0, 0, 8};
  vec sym = {0, 0, 0, 8}, ends = {0, 0, 0, 8};  /* mode 1 scratch: sorted symbols / raw endpoints of one graph */
  PyObject* labels = need_labels ? PyList_New(0) : NULL;
  int eligible = 1, any_weight = 0, failed = 0;
  int32_t base = 0;
  if (need_labels && !labels) failed = 1;
  if (!failed && vec_push(&gptr, &base)) failed = 1;
  for (Py_ssize_t gi = 0; gi < N && eligible && !failed; ++gi) {
    PyObject* el = PySequence_Fast_GET_ITEM(seq, gi);
    if (!(PyTuple_CheckExact(el) || PyList_CheckExact(el)) || PySequence_Fast_GET_SIZE(el) < 2) { eligible = 0; break; }
    PyObject* g = PySequence_Fast_GET_ITEM(el, 0);
    PyObject* L = PySequence_Fast_GET_ITEM(el, 1);
    if (!PyDict_CheckExact(g) || PyDict_Size(g) == 0) { eligible = 0; break; }
    if ((mode == 0 || need_labels) && (!PyDict_CheckExact(L) || PyDict_Size(L) == 0)) { eligible = 0; break; }
    Py_ssize_t pos = 0;
    PyObject *key, *val;
    int64_t n_vert = 0, k0 = 0;
    if (mode == 0) {
      /* label keys must be k0, k0 + 1, ... in key order */
      int64_t expect = 0;
      int first = 1;
      while (PyDict_Next(L, &pos, &key, &val)) {
        int64_t k;
        if (!as_i64(key, &k)) { eligible = 0; break; }
        if (first) { k0 = k; expect = k; first = 0; }
        if (k != expect) { eligible = 0; break; }
        expect += 1;
        if (need_labels && PyList_Append(labels, val)) { failed = 1; break; }
      }
      if (!eligible || failed) break;
      n_vert = expect - k0;
      pos = 0;
      while (PyDict_Next(g, &pos, &key, &val)) {
        int64_t u, v;
        double w;
        if (!PyTuple_CheckExact(key) || PyTuple_GET_SIZE(key) != 2 || !as_i64(PyTuple_GET_ITEM(key, 0), &u) ||
            !as_i64(PyTuple_GET_ITEM(key, 1), &v) || !as_weight(val, &w)) { eligible = 0; break; }
        u -= k0; v -= k0;
        if (u < 0 || u >= n_vert || v < 0 || v >= n_vert) { eligible = 0; break; }  /* unlabelled / foreign vertex */
        const int32_t a = base + (int32_t)u, b = base + (int32_t)v;
        if (vec_push(&src, &a) || vec_push(&dst, &b) || vec_push(&wts, &w)) { failed = 1; break; }
        if (w != 1.0) any_weight = 1;
      }
    } else {
      ends.n = 0;
      const size_t e0 = src.n;
      while (PyDict_Next(g, &pos, &key, &val)) {
        int64_t u, v;
        double w;
        if (!PyTuple_CheckExact(key) || PyTuple_GET_SIZE(key) != 2 || !as_i64(PyTuple_GET_ITEM(key, 0), &u) ||
            !as_i64(PyTuple_GET_ITEM(key, 1), &v) || !as_weight(val, &w)) { eligible = 0; break; }
        if (vec_push(&ends, &u) || vec_push(&ends, &v) || vec_push(&wts, &w)) { failed = 1; break; }
        const int32_t zero = 0;
        if (vec_push(&src, &zero) || vec_push(&dst, &zero)) { failed = 1; break; }
        if (w != 1.0) any_weight = 1;
      }
      if (!eligible || failed) break;
      /* sorted unique symbols */
      sym.n = 0;
      for (size_t i = 0; i < ends.n; ++i)
        if (vec_push(&sym, (int64_t*)ends.p + i)) { failed = 1; break; }
      if (failed) break;
      qsort(sym.p, sym.n, 8, cmp_i64);
      size_t m = 0;
      int64_t* s = (int64_t*)sym.p;
      for (size_t i = 0; i < sym.n; ++i)
        if (i == 0 || s[i] != s[m - 1]) s[m++] = s[i];
      n_vert = (int64_t)m;
      const int64_t* en = (const int64_t*)ends.p;
      for (size_t e = 0; e < ends.n / 2; ++e) {
        const int64_t* pu = (const int64_t*)bsearch(&en[2 * e], s, m, 8, cmp_i64);
        const int64_t* pv = (const int64_t*)bsearch(&en[2 * e + 1], s, m, 8, cmp_i64);
        ((int32_t*)src.p)[e0 + e] = base + (int32_t)(pu - s);
        ((int32_t*)dst.p)[e0 + e] = base + (int32_t)(pv - s);
      }
      if (need_labels) {
        for (size_t i = 0; i < m && eligible && !failed; ++i) {
          PyObject* k = PyLong_FromLongLong(s[i]);
          if (!k) { failed = 1; break; }
          PyObject* lab = PyDict_GetItemWithError(L, k);  /* borrowed */
          Py_DECREF(k);
          if (!lab) {
            if (PyErr_Occurred()) { failed = 1; break; }
            eligible = 0;  /* a vertex of an edge without a label: the Python path raises the reference's KeyError */
            break;
          }
          if (PyList_Append(labels, lab)) { failed = 1; break; }
        }
      }
    }
    if (!eligible || failed) break;
    if ((int64_t)base + n_vert >= INT32_MAX) { eligible = 0; break; }
    base += (int32_t)n_vert;
    if (vec_push(&gptr, &base)) { failed = 1; break; }
  }
  PyObject* result = NULL;
  if (!failed && eligible && N > 0) {
    /* CSR: stable counting sort by source, insertion sort of every (short) row by target */
    const size_t V = (size_t)base, E = src.n;
    int32_t* rp = (int32_t*)calloc(V + 1, 4);
    int32_t* ci = (int32_t*)malloc((E ? E : 1) * 4);
    double* wo = (double*)malloc((E ? E : 1) * 8);
    int32_t* cur = (int32_t*)malloc((V ? V : 1) * 4);
    if (rp && ci && wo && cur) {
      const int32_t* sp = (const int32_t*)src.p;
      const int32_t* dp = (const int32_t*)dst.p;
      const double* wp = (const double*)wts.p;
      for (size_t e = 0; e < E; ++e) rp[sp[e] + 1] += 1;
      for (size_t v = 0; v < V; ++v) rp[v + 1] += rp[v];
      memcpy(cur, rp, V * 4);
      for (size_t e = 0; e < E; ++e) {
        const int32_t k = cur[sp[e]]++;
        ci[k] = dp[e];
        wo[k] = wp[e];
      }
      for (size_t v = 0; v < V; ++v) {
        const int32_t lo = rp[v], hi = rp[v + 1];
        for (int32_t x = lo + 1; x < hi; ++x) {
          const int32_t c = ci[x];
          const double w = wo[x];
          int32_t y = x;
          while (y > lo && ci[y - 1] > c) { ci[y] = ci[y - 1]; wo[y] = wo[y - 1]; --y; }
          ci[y] = c;
          wo[y] = w;
        }
      }
      PyObject* b0 = PyBytes_FromStringAndSize((const char*)gptr.p, (Py_ssize_t)(gptr.n * 4));
      PyObject* b1 = PyBytes_FromStringAndSize((const char*)rp, (Py_ssize_t)((V + 1) * 4));
      PyObject* b2 = PyBytes_FromStringAndSize((const char*)ci, (Py_ssize_t)(E * 4));
      PyObject* b3 = PyBytes_FromStringAndSize((const char*)wo, (Py_ssize_t)(E * 8));
      if (b0 && b1 && b2 && b3)
        result = Py_BuildValue("(OOOOOi)", b0, b1, b2, b3, labels ? labels : Py_None, any_weight);
      Py_XDECREF(b0); Py_XDECREF(b1); Py_XDECREF(b2); Py_XDECREF(b3);
    } else {
      PyErr_NoMemory();
    }
    free(rp); free(ci); free(wo); free(cur);
  } else if (failed && !PyErr_Occurred()) {
    PyErr_NoMemory();
  }
  free(gptr.p); free(src.p); free(dst.p); free(wts.p); free(sym.p); free(ends.p);
  Py_XDECREF(labels);
  Py_DECREF(seq);
  if (result) return result;
  if (PyErr_Occurred()) return NULL;
  Py_RETURN_NONE;
}

static PyMethodDef methods[] = {
    {"pack_edge_dicts", pack_edge_dicts, METH_VARARGS,
     "pack_edge_dicts(elements, mode, need_labels) -> None | (graph_ptr, row_ptr, col_idx, weights, labels, any_weight)"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_fastpack", "CPython-level packer of grakel_b200 (host side)", -1, methods};

PyMODINIT_FUNC PyInit__fastpack(void) { return PyModule_Create(&moddef); }
