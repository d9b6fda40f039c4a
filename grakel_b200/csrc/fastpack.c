/* _fastpack -- CPython-level, multi-threaded packer for the commonest input spelling of the hot path.
 *
 * The reference turns every element [graph, node labels] of a `fit` into a `Graph` object in pure Python
 * (graph.py:147-230, 982-1053: ~1 ms per graph); grakel_b200.packing.pack does the same normalisation with numpy
 * (~0.1 ms per graph).  This module walks the dictionaries with the CPython API directly for the spelling
 * {(u, v): w}  with integer vertex symbols and dictionary labels, and builds the CSR block of ALL graphs in one
 * call.  Anything else -- other spellings, other symbol types, unlabelled or foreign vertices, empty graphs, huge
 * integers -- makes it return None and the Python packer takes over, so semantics and error behaviour stay in one
 * place (packing.py).
 *
 * Threading.  The calling thread keeps the GIL for the whole call, so no Python code runs and no object can
 * change; worker threads (plain pthreads) then READ the dictionaries, tuples, ints and floats of disjoint graph
 * ranges with PyDict_Next and the inline accessor macros only -- nothing that touches reference counts, the
 * allocator or the thread state.  Per-graph sizes are known up front (len of the dictionaries), so every worker
 * writes its graphs straight into the final arrays.
 *
 * pack_edge_dicts(elements, mode, need_labels, want_weights, n_threads)
 *     -> None | (graph_ptr, row_ptr, col_idx, weights | None, labels, any_weight)
 *   elements     list of (graph dict, label dict) pairs
 *   mode         0: vertex set = label keys, a contiguous integer range in key order  (packing mode 'wl')
 *                1: vertex set = the symbols that occur in an edge, sorted            (packing modes 'sp', 'wloa')
 *   graph_ptr / row_ptr / col_idx : bytearray of int32; weights : bytearray of float64 (only if want_weights and a
 *   weight differs from 1); labels : bytearray of int64 when every label is an exact int, else a list of the label
 *   objects, None without need_labels.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <time.h>
static double fp_now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + t.tv_nsec * 1e-9; }

#if PY_VERSION_HEX < 0x030C0000 || defined(Py_GIL_DISABLED)
#define FP_MAX_THREADS 1 /* compact-int accessors / GIL assumptions below need CPython >= 3.12 with the GIL */
#else
#define FP_MAX_THREADS 64
#endif

/* exact int -> int64 without touching the interpreter state; 0 when not an exact small int */
static inline int as_i64(PyObject* o, int64_t* out) {
  if (!PyLong_CheckExact(o)) return 0;
#if PY_VERSION_HEX >= 0x030C0000
  if (!PyUnstable_Long_IsCompact((const PyLongObject*)o)) return 0;
  *out = (int64_t)PyUnstable_Long_CompactValue((const PyLongObject*)o);
  return 1;
#else
  int overflow = 0;
  const long long v = PyLong_AsLongLongAndOverflow(o, &overflow);
  if (overflow || (v == -1 && PyErr_Occurred())) { PyErr_Clear(); return 0; }
  *out = (int64_t)v;
  return 1;
#endif
}

static inline int as_weight(PyObject* o, double* out) {
  if (PyFloat_CheckExact(o)) { *out = PyFloat_AS_DOUBLE(o); return 1; }
  if (o == Py_True) { *out = 1.0; return 1; }
  if (o == Py_False) { *out = 0.0; return 1; }
  int64_t v;
  if (as_i64(o, &v)) { *out = (double)v; return 1; }
  return 0;
}

typedef struct {
  /* inputs */
  PyObject** gd;  /* graph dict of every element */
  PyObject** ld;  /* label dict of every element (may be NULL when unused) */
  Py_ssize_t N;
  int mode, need_labels, want_weights;
  const int64_t* eoff; /* [N + 1] first edge of each graph (= prefix of dict sizes) */
  int64_t* voff;       /* [N + 1] first vertex of each graph; mode 1: filled between the passes */
  int32_t* nvert;      /* [N] mode 1: vertices per graph (pass A) */
  /* outputs */
  int32_t* row_ptr;
  int32_t* col_idx;
  double* weights;
  int64_t* lab_i64;   /* [V] integer label values (valid while all_int) */
  /* scratch, indexed by edge offset */
  int32_t* e_src;
  int32_t* e_dst;
  double* e_w;
  int64_t* sym;       /* mode 1: [2 E] sorted unique symbols of each graph at 2 * eoff[g] */
  /* flags (written with relaxed atomics; only ever set) */
  int eligible, any_weight, all_int;
} job_t;

typedef struct {
  job_t* job;
  Py_ssize_t g0, g1;
  int pass;
  int oom;
} work_t;

static int cmp_i64(const void* a, const void* b) {
  const int64_t x = *(const int64_t*)a, y = *(const int64_t*)b;
  return x < y ? -1 : x > y;
}

/* CSR rows of one graph from its edge list (local vertex ids), rows sorted by target: two stable counting
 * sorts (by target, then by source) -- O(n + m) whatever the degree distribution. */
static int build_rows(job_t* J, int64_t e0, int64_t m, int64_t v0, int64_t n, int32_t* cnt /* [n + 1] scratch */,
                      int32_t* ord /* [m] scratch */) {
  const int32_t* src = J->e_src + e0;
  const int32_t* dst = J->e_dst + e0;
  const double* w = J->e_w ? J->e_w + e0 : NULL;
  memset(cnt, 0, (size_t)(n + 1) * 4);
  for (int64_t e = 0; e < m; ++e) cnt[dst[e] + 1] += 1;
  for (int64_t v = 0; v < n; ++v) cnt[v + 1] += cnt[v];
  for (int64_t e = 0; e < m; ++e) ord[cnt[dst[e]]++] = (int32_t)e;
  memset(cnt, 0, (size_t)(n + 1) * 4);
  for (int64_t e = 0; e < m; ++e) cnt[src[e] + 1] += 1;
  for (int64_t v = 0; v < n; ++v) cnt[v + 1] += cnt[v];
  int32_t* rp = J->row_ptr + v0;
  for (int64_t v = 0; v < n; ++v) rp[v] = (int32_t)(e0 + cnt[v]);
  int32_t* ci = J->col_idx + e0;
  double* wo = (J->weights && w) ? J->weights + e0 : NULL;
  for (int64_t k = 0; k < m; ++k) {
    const int32_t e = ord[k];
    const int32_t at = cnt[src[e]]++;
    ci[at] = (int32_t)(v0 + dst[e]);
    if (wo) wo[at] = w[e];
  }
  return 0;
}

static void* worker(void* arg) {
  work_t* W = (work_t*)arg;
  job_t* J = W->job;
  const double t_w0 = fp_now();
  int32_t* cnt = NULL;
  int32_t* ord = NULL;
  size_t cnt_cap = 0, ord_cap = 0;
  int any_weight = 0;
  for (Py_ssize_t gi = W->g0; gi < W->g1; ++gi) {
    if (!__atomic_load_n(&J->eligible, __ATOMIC_RELAXED)) break;
    PyObject* g = J->gd[gi];
    PyObject* L = J->ld ? J->ld[gi] : NULL;
    const int64_t e0 = J->eoff[gi], m = J->eoff[gi + 1] - e0;
    Py_ssize_t pos = 0;
    PyObject *key, *val;
    int bad = 0;
    if (W->pass == 0 && J->mode == 0) {
      /* ---- 'wl': label keys must be k0, k0 + 1, ... in key order; they are the vertex set */
      const int64_t v0 = J->voff[gi], n = J->voff[gi + 1] - v0;
      int64_t k0 = 0, expect = 0;
      int first = 1;
      while (PyDict_Next(L, &pos, &key, &val)) {
        int64_t k;
        if (!as_i64(key, &k)) { bad = 1; break; }
        if (first) { k0 = k; expect = k; first = 0; }
        if (k != expect) { bad = 1; break; }
        if (J->need_labels) {
          int64_t lv;
          if (as_i64(val, &lv)) J->lab_i64[v0 + (expect - k0)] = lv;
          else __atomic_store_n(&J->all_int, 0, __ATOMIC_RELAXED);
        }
        expect += 1;
      }
      if (bad || expect - k0 != n) { __atomic_store_n(&J->eligible, 0, __ATOMIC_RELAXED); break; }
      pos = 0;
      int64_t e = 0;
      while (PyDict_Next(g, &pos, &key, &val)) {
        int64_t u, v;
        double w;
        if (e >= m || !PyTuple_CheckExact(key) || PyTuple_GET_SIZE(key) != 2 || !as_i64(PyTuple_GET_ITEM(key, 0), &u) ||
            !as_i64(PyTuple_GET_ITEM(key, 1), &v) || !as_weight(val, &w)) { bad = 1; break; }
        u -= k0; v -= k0;
        if (u < 0 || u >= n || v < 0 || v >= n) { bad = 1; break; }  /* unlabelled / foreign vertex */
        J->e_src[e0 + e] = (int32_t)u;
        J->e_dst[e0 + e] = (int32_t)v;
        if (J->e_w) J->e_w[e0 + e] = w;
        if (w != 1.0) any_weight = 1;
        ++e;
      }
      if (bad || e != m) { __atomic_store_n(&J->eligible, 0, __ATOMIC_RELAXED); break; }
      if ((size_t)(n + 1) > cnt_cap) { free(cnt); cnt_cap = (size_t)(n + 1) * 2; cnt = (int32_t*)malloc(cnt_cap * 4); }
      if ((size_t)m > ord_cap) { free(ord); ord_cap = (size_t)m * 2; ord = (int32_t*)malloc(ord_cap * 4); }
      if (!cnt || (m && !ord)) { W->oom = 1; break; }
      build_rows(J, e0, m, v0, n, cnt, ord);
    } else if (W->pass == 0) {
      /* ---- 'sp' / 'wloa', pass A: endpoints -> sorted unique symbols -> local indices */
      int64_t* s = J->sym + 2 * e0;
      int64_t e = 0;
      while (PyDict_Next(g, &pos, &key, &val)) {
        int64_t u, v;
        double w;
        if (e >= m || !PyTuple_CheckExact(key) || PyTuple_GET_SIZE(key) != 2 || !as_i64(PyTuple_GET_ITEM(key, 0), &u) ||
            !as_i64(PyTuple_GET_ITEM(key, 1), &v) || !as_weight(val, &w)) { bad = 1; break; }
        s[2 * e] = u; s[2 * e + 1] = v;
        if (J->e_w) J->e_w[e0 + e] = w;
        if (w != 1.0) any_weight = 1;
        ++e;
      }
      if (bad || e != m) { __atomic_store_n(&J->eligible, 0, __ATOMIC_RELAXED); break; }
      /* e_src / e_dst: positions of the endpoints in the sorted symbol list */
      int64_t* raw = (int64_t*)malloc((size_t)(2 * m) * 8 + 8);
      if (!raw) { W->oom = 1; break; }
      memcpy(raw, s, (size_t)(2 * m) * 8);
      qsort(s, (size_t)(2 * m), 8, cmp_i64);
      int64_t nu = 0;
      for (int64_t i = 0; i < 2 * m; ++i)
        if (i == 0 || s[i] != s[nu - 1]) s[nu++] = s[i];
      for (int64_t k = 0; k < m; ++k) {
        const int64_t* pu = (const int64_t*)bsearch(&raw[2 * k], s, (size_t)nu, 8, cmp_i64);
        const int64_t* pv = (const int64_t*)bsearch(&raw[2 * k + 1], s, (size_t)nu, 8, cmp_i64);
        J->e_src[e0 + k] = (int32_t)(pu - s);
        J->e_dst[e0 + k] = (int32_t)(pv - s);
      }
      free(raw);
      J->nvert[gi] = (int32_t)nu;
    } else {
      /* ---- 'sp' / 'wloa', pass B: rows at the final offsets, labels of the sorted symbols */
      const int64_t v0 = J->voff[gi], n = J->voff[gi + 1] - v0;
      if ((size_t)(n + 1) > cnt_cap) { free(cnt); cnt_cap = (size_t)(n + 1) * 2; cnt = (int32_t*)malloc(cnt_cap * 4); }
      if ((size_t)m > ord_cap) { free(ord); ord_cap = (size_t)m * 2; ord = (int32_t*)malloc(ord_cap * 4); }
      if (!cnt || (m && !ord)) { W->oom = 1; break; }
      build_rows(J, e0, m, v0, n, cnt, ord);
      if (J->need_labels) {
        /* walk L once: every symbol of the graph must be a key; other keys are ignored (graph.py:390-394) */
        const int64_t* s = J->sym + 2 * e0;
        int64_t found = 0;
        while (PyDict_Next(L, &pos, &key, &val)) {
          int64_t k;
          if (!as_i64(key, &k)) { bad = 1; break; }  /* non-int key: let the Python path decide */
          const int64_t* pk = (const int64_t*)bsearch(&k, s, (size_t)n, 8, cmp_i64);
          if (!pk) continue;
          int64_t lv;
          if (as_i64(val, &lv)) J->lab_i64[v0 + (pk - s)] = lv;
          else __atomic_store_n(&J->all_int, 0, __ATOMIC_RELAXED);
          ++found;
        }
        if (bad || found != n) { __atomic_store_n(&J->eligible, 0, __ATOMIC_RELAXED); break; }  /* unlabelled vertex: KeyError path */
      }
    }
  }
  if (any_weight) __atomic_store_n(&J->any_weight, 1, __ATOMIC_RELAXED);
  if (getenv("GRAKEL_B200_PACK_DEBUG")) fprintf(stderr, "  [worker] graphs %ld..%ld pass %d: %.2f ms\n", (long)W->g0, (long)W->g1, W->pass, (fp_now() - t_w0) * 1e3);
  free(cnt);
  free(ord);
  return NULL;
}

/* run one pass over all graphs on nt threads (the caller is one of them); graphs are split by edge count */
static int run_pass(job_t* J, int pass, int nt) {
  work_t wk[FP_MAX_THREADS];
  pthread_t th[FP_MAX_THREADS];
  const int64_t E = J->eoff[J->N];
  Py_ssize_t g = 0;
  for (int t = 0; t < nt; ++t) {
    wk[t].job = J; wk[t].pass = pass; wk[t].oom = 0;
    wk[t].g0 = g;
    const int64_t want = E * (t + 1) / nt;
    while (g < J->N && (t == nt - 1 || J->eoff[g + 1] <= want)) ++g;
    wk[t].g1 = g;
  }
  wk[nt - 1].g1 = J->N;
  int started = 0;
  for (int t = 1; t < nt; ++t) {
    if (pthread_create(&th[t], NULL, worker, &wk[t]) != 0) break;
    ++started;
  }
  worker(&wk[0]);
  for (int t = started + 1; t < nt; ++t) worker(&wk[t]);  /* threads that could not be created: run inline */
  for (int t = 1; t <= started; ++t) pthread_join(th[t], NULL);
  for (int t = 0; t < nt; ++t)
    if (wk[t].oom) return -1;
  return 0;
}

static PyObject* pack_edge_dicts(PyObject* self, PyObject* args) {
  PyObject* seq_in;
  int mode, need_labels, want_weights = 1, n_threads = 1;
  if (!PyArg_ParseTuple(args, "Oii|ii", &seq_in, &mode, &need_labels, &want_weights, &n_threads)) return NULL;
  PyObject* seq = PySequence_Fast(seq_in, "elements must be a sequence");
  if (!seq) return NULL;
  const Py_ssize_t N = PySequence_Fast_GET_SIZE(seq);
  if (N == 0) { Py_DECREF(seq); Py_RETURN_NONE; }
  if (n_threads < 1) n_threads = 1;
  if (n_threads > FP_MAX_THREADS) n_threads = FP_MAX_THREADS;

  const int fp_debug = getenv("GRAKEL_B200_PACK_DEBUG") != NULL;
  const double t_start = fp_now();
  double t_pre = 0, t_alloc = 0, t_passA = 0, t_passB = 0;
  job_t J;
  memset(&J, 0, sizeof(J));
  J.N = N; J.mode = mode; J.need_labels = need_labels; J.want_weights = want_weights;
  J.eligible = 1; J.all_int = 1;
  const int use_labels = (mode == 0 || need_labels);
  PyObject** gd = (PyObject**)malloc((size_t)N * sizeof(PyObject*));
  PyObject** ld = use_labels ? (PyObject**)malloc((size_t)N * sizeof(PyObject*)) : NULL;
  int64_t* eoff = (int64_t*)malloc((size_t)(N + 1) * 8);
  int64_t* voff = (int64_t*)malloc((size_t)(N + 1) * 8);
  int32_t* nvert = (int32_t*)calloc((size_t)N, 4);
  PyObject *b_gp = NULL, *b_rp = NULL, *b_ci = NULL, *b_w = NULL, *b_lab = NULL, *lab_list = NULL, *result = NULL;
  int eligible = 1, failed = 0;
  if (!gd || (use_labels && !ld) || !eoff || !voff || !nvert) { failed = 1; goto done; }

  /* ---- serial pre-pass (GIL): element shapes and sizes */
  eoff[0] = 0; voff[0] = 0;
  for (Py_ssize_t gi = 0; gi < N; ++gi) {
    PyObject* el = PySequence_Fast_GET_ITEM(seq, gi);
    if (!(PyTuple_CheckExact(el) || PyList_CheckExact(el)) || PySequence_Fast_GET_SIZE(el) < 2) { eligible = 0; break; }
    PyObject* g = PySequence_Fast_GET_ITEM(el, 0);
    PyObject* L = PySequence_Fast_GET_ITEM(el, 1);
    if (!PyDict_CheckExact(g) || PyDict_GET_SIZE(g) == 0) { eligible = 0; break; }
    if (use_labels && (!PyDict_CheckExact(L) || PyDict_GET_SIZE(L) == 0)) { eligible = 0; break; }
    gd[gi] = g;
    if (ld) ld[gi] = L;
    eoff[gi + 1] = eoff[gi] + (int64_t)PyDict_GET_SIZE(g);
    voff[gi + 1] = voff[gi] + (mode == 0 ? (int64_t)PyDict_GET_SIZE(L) : 0);
  }
  if (!eligible) goto done;
  t_pre = fp_now();
  {
    const int64_t E = eoff[N];
    if (E >= INT32_MAX || voff[N] >= INT32_MAX) { eligible = 0; goto done; }  /* the general path raises the int32 error */
    J.gd = gd; J.ld = ld; J.eoff = eoff; J.voff = voff; J.nvert = nvert;
    J.e_src = (int32_t*)malloc((size_t)(E ? E : 1) * 4);
    J.e_dst = (int32_t*)malloc((size_t)(E ? E : 1) * 4);
    J.e_w = want_weights ? (double*)malloc((size_t)(E ? E : 1) * 8) : NULL;
    J.sym = mode == 1 ? (int64_t*)malloc((size_t)(E ? E : 1) * 16) : NULL;
    b_ci = PyByteArray_FromStringAndSize(NULL, (Py_ssize_t)(E * 4));
    if (want_weights) b_w = PyByteArray_FromStringAndSize(NULL, (Py_ssize_t)(E * 8));
    if (!J.e_src || !J.e_dst || (want_weights && (!J.e_w || !b_w)) || (mode == 1 && !J.sym) || !b_ci) { failed = 1; goto done; }
    J.col_idx = (int32_t*)PyByteArray_AS_STRING(b_ci);
    J.weights = b_w ? (double*)PyByteArray_AS_STRING(b_w) : NULL;
    int nt = n_threads;
    if (E < 20000) nt = 1;  /* thread start-up is not worth it */
    if (nt > N) nt = (int)N;

    t_alloc = fp_now();
    if (mode == 1) {
      if (run_pass(&J, 0, nt)) { failed = 1; goto done; }
      if (!J.eligible) { eligible = 0; goto done; }
      for (Py_ssize_t gi = 0; gi < N; ++gi) voff[gi + 1] = voff[gi] + nvert[gi];
      if (voff[N] >= INT32_MAX) { eligible = 0; goto done; }
    }
    t_passA = fp_now();
    const int64_t V = voff[N];
    b_gp = PyByteArray_FromStringAndSize(NULL, (Py_ssize_t)((N + 1) * 4));
    b_rp = PyByteArray_FromStringAndSize(NULL, (Py_ssize_t)((V + 1) * 4));
    if (need_labels) b_lab = PyByteArray_FromStringAndSize(NULL, (Py_ssize_t)(V * 8));
    if (!b_gp || !b_rp || (need_labels && !b_lab)) { failed = 1; goto done; }
    J.row_ptr = (int32_t*)PyByteArray_AS_STRING(b_rp);
    J.lab_i64 = b_lab ? (int64_t*)PyByteArray_AS_STRING(b_lab) : NULL;
    if (run_pass(&J, mode == 1 ? 1 : 0, nt)) { failed = 1; goto done; }
    if (!J.eligible) { eligible = 0; goto done; }
    t_passB = fp_now();
    J.row_ptr[V] = (int32_t)E;
    int32_t* gp = (int32_t*)PyByteArray_AS_STRING(b_gp);
    for (Py_ssize_t gi = 0; gi <= N; ++gi) gp[gi] = (int32_t)voff[gi];

    PyObject* labels_out = Py_None;
    if (need_labels) {
      if (J.all_int) {
        labels_out = b_lab;
      } else {
        /* general label objects: collected serially under the GIL, in vertex order */
        lab_list = PyList_New((Py_ssize_t)V);
        if (!lab_list) { failed = 1; goto done; }
        for (Py_ssize_t gi = 0; gi < N && !failed; ++gi) {
          PyObject* L = ld[gi];
          const int64_t v0 = voff[gi], n = voff[gi + 1] - v0;
          Py_ssize_t pos = 0;
          PyObject *key, *val;
          if (mode == 0) {
            int64_t i = 0;
            while (PyDict_Next(L, &pos, &key, &val) && i < n) {
              Py_INCREF(val);
              PyList_SET_ITEM(lab_list, (Py_ssize_t)(v0 + i), val);
              ++i;
            }
          } else {
            const int64_t* s = J.sym + 2 * eoff[gi];
            while (PyDict_Next(L, &pos, &key, &val)) {
              int64_t k;
              if (!as_i64(key, &k)) continue;
              const int64_t* pk = (const int64_t*)bsearch(&k, s, (size_t)n, 8, cmp_i64);
              if (!pk) continue;
              Py_INCREF(val);
              PyList_SET_ITEM(lab_list, (Py_ssize_t)(v0 + (pk - s)), val);
            }
          }
        }
        labels_out = lab_list;
      }
    }
    PyObject* w_out = (want_weights && J.any_weight) ? b_w : Py_None;
    result = Py_BuildValue("(OOOOOi)", b_gp, b_rp, b_ci, w_out, labels_out, J.any_weight);
    if (fp_debug)
      fprintf(stderr, "[fastpack] threads %d: pre-pass %.2f ms, alloc %.2f, pass A %.2f, pass B (incl. output alloc) %.2f, labels/result %.2f\n", nt,
              (t_pre - t_start) * 1e3, (t_alloc - t_pre) * 1e3, (t_passA - t_alloc) * 1e3, (t_passB - t_passA) * 1e3, (fp_now() - t_passB) * 1e3);
  }
done:
  free(gd); free(ld); free(eoff); free(voff); free(nvert);
  free(J.e_src); free(J.e_dst); free(J.e_w); free(J.sym);
  Py_XDECREF(b_gp); Py_XDECREF(b_rp); Py_XDECREF(b_ci); Py_XDECREF(b_w); Py_XDECREF(b_lab); Py_XDECREF(lab_list);
  Py_DECREF(seq);
  if (result) return result;
  if (failed && !PyErr_Occurred()) PyErr_NoMemory();
  if (PyErr_Occurred()) return NULL;
  Py_RETURN_NONE;
}

static PyMethodDef methods[] = {
    {"pack_edge_dicts", pack_edge_dicts, METH_VARARGS,
     "pack_edge_dicts(elements, mode, need_labels, want_weights=1, n_threads=1) -> None | (graph_ptr, row_ptr, col_idx, "
     "weights | None, labels, any_weight)"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_fastpack", "CPython-level packer of grakel_b200 (host side)", -1, methods};

PyMODINIT_FUNC PyInit__fastpack(void) { return PyModule_Create(&moddef); }
