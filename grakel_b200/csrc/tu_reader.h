// TU-format dataset files -> packed CSR block, host side (no CUDA).
//
// Replaces, for the hot path's input, grakel/datasets/base.py:135-290 (`read_data`: five text
// files parsed line by line into per-graph Python sets / dicts) followed by the kernels' own
// Graph parsing (graph.py:147-230, 982-1053): on 10 000 graphs that pair costs seconds of Python
// per call while the device pass takes under a millisecond.  Files:
//   <dir>/<name>_graph_indicator.txt   line i: graph id (1-based) of node i
//   <dir>/<name>_A.txt                 line e: "u, v" (1-based node ids)
//   <dir>/<name>_node_labels.txt       line i: integer label of node i            (optional)
//   <dir>/<name>_node_attributes.txt   line i: comma separated floats             (optional)
//   <dir>/<name>_edge_labels.txt       line e: integer label of edge line e       (optional)
//   <dir>/<name>_graph_labels.txt      line g: class of graph g                   (optional)
// Semantics kept from the reference: an edge line belongs to the graph of its source; with
// `is_symmetric` the reverse edge is added to the graph of the target (base.py:216-218); edges are
// SETS (duplicates collapse), a later edge-label line overwrites an earlier one (:262-266).
#pragma once
#include <algorithm>
#include <cerrno>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../../include/grakel_b200.h"

struct gk_tu {
  int64_t n_graphs = 0, n_nodes = 0;
  int32_t flags = 0;
  std::vector<int32_t> indicator;         // node -> graph (0-based)
  std::vector<int32_t> eu, ev, el;        // edge lines (0-based node ids); el empty without edge labels
  std::vector<int32_t> node_label;        // empty without node labels
  std::vector<double> node_attr;          // [n_nodes * attr_dim]
  int32_t attr_dim = 0;
  std::vector<int32_t> classes;
  // packed by gk_tu_pack
  int32_t mode = -1;
  std::vector<int32_t> graph_ptr, row_ptr, col_idx, vnode, elabel, vlabel;
};

namespace gk {
namespace tu {

struct FileBuf {
  std::unique_ptr<char[]> buf;  // not zero-filled: every byte is overwritten by fread / the sentinel
  size_t size = 0;              // file bytes + the sentinel newline
  bool present = false;
  const char* begin() const { return buf.get(); }
};

inline bool slurp(const std::string& path, FileBuf* fb, std::string* err) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) {
    if (errno == ENOENT) return true;  // optional file
    *err = "cannot open " + path + ": " + strerror(errno);
    return false;
  }
  fseek(f, 0, SEEK_END);
  const long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  fb->size = (size_t)std::max<long>(n, 0) + 1;
  fb->buf.reset(new char[fb->size]);
  const size_t got = n > 0 ? fread(fb->buf.get(), 1, (size_t)n, f) : 0;
  fclose(f);
  if ((long)got != n) { *err = "short read on " + path; return false; }
  fb->buf[(size_t)n] = '\n';  // sentinel: every line ends
  fb->present = true;
  return true;
}

// one signed integer per field; fields separated by ',' and blanks; `per_line` fields per line.
// [p, end) must end with a newline (the file sentinel or a chunk boundary).
inline bool parse_int_range(const char* p, const char* end, int per_line, std::vector<int32_t>* out0,
                            std::vector<int32_t>* out1, const std::string& what, std::string* err) {
  long line = 0;
  while (p < end) {
    ++line;
    int got = 0;
    long long vals[2] = {0, 0};
    while (p < end && *p != '\n') {
      while (p < end && (*p == ' ' || *p == '\t' || *p == ',' || *p == '\r')) ++p;
      if (p >= end || *p == '\n') break;
      bool neg = false;
      if (*p == '-' || *p == '+') { neg = *p == '-'; ++p; }
      if (*p < '0' || *p > '9') { *err = what + ": line " + std::to_string(line) + ": not an integer"; return false; }
      long long v = 0;
      while (*p >= '0' && *p <= '9') {
        v = v * 10 + (*p - '0');
        if (v > (1LL << 40)) { *err = what + ": line " + std::to_string(line) + ": integer out of range"; return false; }
        ++p;
      }
      if (got < 2) vals[got] = neg ? -v : v;
      ++got;
    }
    if (p < end) ++p;  // newline
    if (got == 0) {
      if (p >= end) break;  // trailing blank line
      *err = what + ": line " + std::to_string(line) + " is empty";
      return false;
    }
    if (got != per_line) { *err = what + ": line " + std::to_string(line) + ": expected " + std::to_string(per_line) + " field(s)"; return false; }
    for (int k = 0; k < per_line; ++k)
      if (vals[k] > INT32_MAX || vals[k] < INT32_MIN) { *err = what + ": line " + std::to_string(line) + ": integer out of range"; return false; }
    out0->push_back((int32_t)vals[0]);
    if (per_line == 2) out1->push_back((int32_t)vals[1]);
  }
  return true;
}

// Output vectors are reserved from the file size (first-touch page faults, not parsing, dominate otherwise).
inline bool parse_ints(const FileBuf& fb, int per_line, std::vector<int32_t>* out0, std::vector<int32_t>* out1,
                       const std::string& what, std::string* err) {
  const size_t est = fb.size / (per_line == 2 ? 8 : 2) + 16;
  out0->reserve(est);
  if (per_line == 2) out1->reserve(est);
  return parse_int_range(fb.begin(), fb.begin() + fb.size, per_line, out0, out1, what, err);
}

inline bool parse_floats(const FileBuf& fb, std::vector<double>* out, int32_t* dim, const std::string& what, std::string* err) {
  const char* p = fb.begin();
  const char* end = p + fb.size - 1;
  long line = 0;
  *dim = 0;
  while (p < end) {
    ++line;
    int got = 0;
    while (p < end && *p != '\n') {
      while (p < end && (*p == ' ' || *p == '\t' || *p == ',' || *p == '\r')) ++p;
      if (p >= end || *p == '\n') break;
      char* q = nullptr;
      const double v = strtod(p, &q);  // the sentinel newline stops it
      if (q == p) { *err = what + ": line " + std::to_string(line) + ": not a number"; return false; }
      p = q;
      out->push_back(v);
      ++got;
    }
    if (p < end) ++p;
    if (got == 0) {
      if (p >= end) break;
      *err = what + ": line " + std::to_string(line) + " is empty";
      return false;
    }
    if (*dim == 0) *dim = got;
    else if (got != *dim) { *err = what + ": line " + std::to_string(line) + ": ragged attribute rows"; return false; }
  }
  return true;
}

inline int open(const char* dir, const char* name, int32_t flags, gk_tu** out, std::string* err) {
  const std::string base = std::string(dir) + "/" + name + "_";
  gk_tu* t = new gk_tu();
  t->flags = flags;
  // the edge list (by far the largest file) is read and parsed on the calling thread, everything else on a helper
  bool ok_side = true, has_indicator = false;
  std::string err_side;
  std::thread side([&]() {
    FileBuf fi, fn, fna, fe, fg;
    std::vector<int32_t> dummy;
    bool ok = slurp(base + "graph_indicator.txt", &fi, &err_side) && slurp(base + "node_labels.txt", &fn, &err_side) &&
              slurp(base + "edge_labels.txt", &fe, &err_side) && slurp(base + "graph_labels.txt", &fg, &err_side);
    if (ok && (flags & GK_TU_ATTR_NODES)) ok = slurp(base + "node_attributes.txt", &fna, &err_side);
    has_indicator = fi.present;
    if (ok && fi.present) ok = parse_ints(fi, 1, &t->indicator, &dummy, "graph_indicator", &err_side);
    // read_data only opens the label file when it does not take the attributes (base.py:223-240)
    if (ok && fn.present && !fna.present) ok = parse_ints(fn, 1, &t->node_label, &dummy, "node_labels", &err_side);
    if (ok && fe.present) ok = parse_ints(fe, 1, &t->el, &dummy, "edge_labels", &err_side);
    if (ok && fg.present) ok = parse_ints(fg, 1, &t->classes, &dummy, "graph_labels", &err_side);
    if (ok && fna.present) ok = parse_floats(fna, &t->node_attr, &t->attr_dim, "node_attributes", &err_side);
    ok_side = ok;
  });
  FileBuf fa;
  bool ok = slurp(base + "A.txt", &fa, err);
  if (ok && fa.present) ok = parse_ints(fa, 2, &t->eu, &t->ev, "A", err);
  fa.buf.reset();
  side.join();
  if (ok_side && !has_indicator) { err_side = "missing " + base + "graph_indicator.txt"; ok_side = false; }
  if (!ok_side) { *err = err_side; ok = false; }  // the indicator / label files are reported first, like a sequential read
  if (ok && !fa.present) { *err = "missing " + base + "A.txt"; ok = false; }
  if (ok) {
    t->n_nodes = (int64_t)t->indicator.size();
    int32_t gmax = 0;
    for (auto& g : t->indicator) {
      if (g < 1) { *err = "graph_indicator: graph ids are 1-based"; ok = false; break; }
      gmax = std::max(gmax, g);
      g -= 1;
    }
    t->n_graphs = gmax;
  }
  if (ok && !t->node_label.empty() && (int64_t)t->node_label.size() != t->n_nodes) { *err = "node_labels: one line per node expected"; ok = false; }
  if (ok && t->attr_dim && (int64_t)t->node_attr.size() != t->n_nodes * t->attr_dim) { *err = "node_attributes: one line per node expected"; ok = false; }
  if (ok && !t->el.empty() && t->el.size() != t->eu.size()) { *err = "edge_labels: one line per edge line expected"; ok = false; }
  if (ok && !t->classes.empty() && (int64_t)t->classes.size() != t->n_graphs) { *err = "graph_labels: one line per graph expected"; ok = false; }
  if (ok) {
    for (size_t e = 0; e < t->eu.size(); ++e) {
      if (t->eu[e] < 1 || t->ev[e] < 1 || t->eu[e] > t->n_nodes || t->ev[e] > t->n_nodes) { *err = "A: node id out of range on line " + std::to_string(e + 1); ok = false; break; }
      t->eu[e] -= 1; t->ev[e] -= 1;
    }
  }
  if (!ok) { delete t; return GK_ERR_ARG; }
  *out = t;
  return GK_OK;
}

// mode GK_TU_LABELLED_NODES: vertex set = the keys of the node-label dictionary (what WeisfeilerLehman walks,
//   weisfeiler_lehman.py:234) -- every node of the graph when a label / attribute file exists;
// mode GK_TU_EDGE_NODES: vertex set = the nodes that occur in an edge of their graph (ShortestPath's
//   sorted edge symbols, graph.py:1613-1631; the edge-dictionary keys WL-OA walks).
//
// Edges are bucketed by source with a STABLE counting sort in the reference's write order (line e forward, then
// its reverse when symmetric) and each short row is sorted by target with a stable insertion sort, so the last
// element of a run of equal targets is the last write of that pair: duplicates collapse and the last edge label
// wins (base.py:262-266) without carrying sequence numbers around.
inline int pack(gk_tu* t, int32_t mode, std::string* err) {
  if (mode != GK_TU_LABELLED_NODES && mode != GK_TU_EDGE_NODES) { *err = "gk_tu_pack: unknown mode"; return GK_ERR_ARG; }
  const int64_t n = t->n_nodes;
  const bool sym = t->flags & GK_TU_SYMMETRIC;
  const bool has_el = !t->el.empty();
  const size_t L = t->eu.size();
  const int32_t* ind = t->indicator.data();
  std::vector<uint32_t> start(n + 1, 0);
  for (size_t e = 0; e < L; ++e) {
    const int32_t u = t->eu[e], v = t->ev[e];
    if (ind[u] != ind[v]) {
      *err = "A: line " + std::to_string(e + 1) + " joins nodes of different graphs (the reference raises KeyError on it)";
      return GK_ERR_ARG;
    }
    start[u + 1] += 1;
    if (sym) start[v + 1] += 1;
  }
  for (int64_t i = 0; i < n; ++i) start[i + 1] += start[i];
  const size_t M = start[n];
  std::unique_ptr<int32_t[]> rv(new int32_t[M ? M : 1]), rl(has_el ? new int32_t[M ? M : 1] : nullptr);
  {
    std::vector<uint32_t> cur(start.begin(), start.end() - 1);
    for (size_t e = 0; e < L; ++e) {
      const int32_t u = t->eu[e], v = t->ev[e];
      uint32_t k = cur[u]++;
      rv[k] = v;
      if (has_el) rl[k] = t->el[e];
      if (sym) {
        k = cur[v]++;
        rv[k] = u;
        if (has_el) rl[k] = t->el[e];
      }
    }
  }
  // sort every row by target (stable), then keep the last element of each run of equal targets
  std::vector<uint32_t> deg(n, 0);
  std::vector<std::pair<int32_t, int32_t>> tmp;
  for (int64_t i = 0; i < n; ++i) {
    const uint32_t lo = start[i], hi = start[i + 1];
    if (hi - lo > 48) {
      tmp.resize(hi - lo);
      for (uint32_t x = lo; x < hi; ++x) tmp[x - lo] = {rv[x], has_el ? rl[x] : 0};
      std::stable_sort(tmp.begin(), tmp.end(), [](const std::pair<int32_t, int32_t>& a, const std::pair<int32_t, int32_t>& b) { return a.first < b.first; });
      for (uint32_t x = lo; x < hi; ++x) { rv[x] = tmp[x - lo].first; if (has_el) rl[x] = tmp[x - lo].second; }
    } else {
      for (uint32_t x = lo + 1; x < hi; ++x) {
        const int32_t v = rv[x], l = has_el ? rl[x] : 0;
        uint32_t y = x;
        while (y > lo && rv[y - 1] > v) { rv[y] = rv[y - 1]; if (has_el) rl[y] = rl[y - 1]; --y; }
        rv[y] = v;
        if (has_el) rl[y] = l;
      }
    }
    uint32_t w = lo;
    for (uint32_t x = lo; x < hi; ++x) {
      if (x + 1 < hi && rv[x + 1] == rv[x]) continue;  // a later write of the same pair follows
      rv[w] = rv[x];
      if (has_el) rl[w] = rl[x];
      ++w;
    }
    deg[i] = w - lo;
  }
  // node labels: file, or (GK_TU_DEGREE_LABELS, no file) out-degree without self loops -- nodes of degree 0 stay
  // unlabelled (base.py:239-241: Counter over the sources of the edge set)
  std::vector<int32_t> label;
  std::vector<char> labelled(n, 0);
  const bool attrs = t->attr_dim > 0;
  if (!t->node_label.empty() && !attrs) {
    label = t->node_label;
    std::fill(labelled.begin(), labelled.end(), 1);
  } else if (attrs) {
    std::fill(labelled.begin(), labelled.end(), 1);
  } else if (t->flags & GK_TU_DEGREE_LABELS) {
    label.assign(n, 0);
    for (int64_t i = 0; i < n; ++i)
      for (uint32_t x = start[i]; x < start[i] + deg[i]; ++x)
        if (rv[x] != (int32_t)i) { label[i] += 1; labelled[i] = 1; }
  }
  std::vector<char> keep(n, 0);
  if (mode == GK_TU_LABELLED_NODES) {
    keep = labelled;
  } else {
    for (int64_t i = 0; i < n; ++i)
      for (uint32_t x = start[i]; x < start[i] + deg[i]; ++x) { keep[i] = 1; keep[rv[x]] = 1; }
    // a vertex of an edge without a label (e.g. degree labels and a node that is only ever an edge target):
    // read_data leaves it unlabelled and ShortestPath raises KeyError on it -- no silent label 0 here
    if (!label.empty())
      for (int64_t i = 0; i < n; ++i)
        if (keep[i] && !labelled[i]) {
          *err = "node " + std::to_string(i + 1) + " occurs in an edge but has no label (the reference raises KeyError)";
          return GK_ERR_ARG;
        }
  }
  // vertices grouped by graph, node order inside a graph (stable counting sort)
  t->graph_ptr.assign(t->n_graphs + 1, 0);
  for (int64_t i = 0; i < n; ++i)
    if (keep[i]) t->graph_ptr[ind[i] + 1] += 1;
  for (int64_t g = 0; g < t->n_graphs; ++g) t->graph_ptr[g + 1] += t->graph_ptr[g];
  const int64_t V = t->graph_ptr[t->n_graphs];
  std::vector<int32_t> pos(n, -1), cursor(t->graph_ptr.begin(), t->graph_ptr.end() - 1);
  t->vnode.assign(V, 0);
  for (int64_t i = 0; i < n; ++i)
    if (keep[i]) { pos[i] = cursor[ind[i]]++; t->vnode[pos[i]] = (int32_t)i; }
  // CSR over the kept vertices.  An edge whose source is not a vertex is never visited by the reference
  // (weisfeiler_lehman.py:234 walks labelled vertices only); a target that is not a vertex is its KeyError.
  t->row_ptr.assign(V + 1, 0);
  for (int64_t i = 0; i < n; ++i) {
    if (pos[i] < 0) continue;
    for (uint32_t x = start[i]; x < start[i] + deg[i]; ++x)
      if (pos[rv[x]] < 0) {
        *err = "node " + std::to_string(rv[x] + 1) + " is the target of an edge but has no label (the reference raises KeyError)";
        return GK_ERR_ARG;
      }
    t->row_ptr[pos[i] + 1] = (int32_t)deg[i];
  }
  for (int64_t v = 0; v < V; ++v) t->row_ptr[v + 1] += t->row_ptr[v];
  const size_t kept_edges = (size_t)t->row_ptr[V];
  t->col_idx.assign(kept_edges, 0);
  t->elabel.assign(has_el ? kept_edges : 0, 0);
  // positions are monotone in the node id inside a graph, so rows sorted by node id stay sorted
  for (int64_t v = 0; v < V; ++v) {
    const int32_t i = t->vnode[v];
    int32_t k = t->row_ptr[v];
    for (uint32_t x = start[i]; x < start[i] + deg[i]; ++x, ++k) {
      t->col_idx[k] = pos[rv[x]];
      if (has_el) t->elabel[k] = rl[x];
    }
  }
  t->vlabel.clear();
  if (!label.empty()) {
    t->vlabel.resize(V);
    for (int64_t v = 0; v < V; ++v) t->vlabel[v] = label[t->vnode[v]];
  }
  t->mode = mode;
  return GK_OK;
}

}  // namespace tu
}  // namespace gk
