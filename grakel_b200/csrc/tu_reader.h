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
  std::vector<char> data;
  bool present = false;
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
  fb->data.resize((size_t)std::max<long>(n, 0) + 1);
  const size_t got = n > 0 ? fread(fb->data.data(), 1, (size_t)n, f) : 0;
  fclose(f);
  if ((long)got != n) { *err = "short read on " + path; return false; }
  fb->data[(size_t)n] = '\n';  // sentinel: every line ends
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

// Large files are cut at newlines and parsed by a few threads; on any error the file is parsed again
// sequentially so that the message carries the exact line number.
inline bool parse_ints(const FileBuf& fb, int per_line, std::vector<int32_t>* out0, std::vector<int32_t>* out1,
                       const std::string& what, std::string* err) {
  const char* base = fb.data.data();
  const size_t size = fb.data.size();  // includes the sentinel newline
  const size_t est = size / (per_line == 2 ? 10 : 3) + 16;
  unsigned T = std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
  if (size < (1u << 20)) T = 1;
  if (T > 1) {
    std::vector<size_t> cut(T + 1, 0);
    cut[T] = size;
    for (unsigned i = 1; i < T; ++i) {
      const char* q = (const char*)memchr(base + size * i / T, '\n', size - size * i / T);
      cut[i] = q ? (size_t)(q - base) + 1 : size;
    }
    std::vector<std::vector<int32_t>> a(T), b(T);
    std::vector<std::string> errs(T);
    std::vector<char> oks(T, 1);
    std::vector<std::thread> th;
    for (unsigned i = 0; i < T; ++i)
      th.emplace_back([&, i]() {
        a[i].reserve(est / T);
        if (per_line == 2) b[i].reserve(est / T);
        if (cut[i] < cut[i + 1]) oks[i] = parse_int_range(base + cut[i], base + cut[i + 1], per_line, &a[i], &b[i], what, &errs[i]);
      });
    for (auto& t : th) t.join();
    bool ok = true;
    for (unsigned i = 0; i < T; ++i) ok = ok && oks[i];
    if (ok) {
      for (unsigned i = 0; i < T; ++i) {
        out0->insert(out0->end(), a[i].begin(), a[i].end());
        if (per_line == 2) out1->insert(out1->end(), b[i].begin(), b[i].end());
      }
      return true;
    }
  }
  out0->reserve(est);
  if (per_line == 2) out1->reserve(est);
  return parse_int_range(base, base + size, per_line, out0, out1, what, err);
}

inline bool parse_floats(const FileBuf& fb, std::vector<double>* out, int32_t* dim, const std::string& what, std::string* err) {
  const char* p = fb.data.data();
  const char* end = p + fb.data.size() - 1;
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
  FileBuf fi, fa, fn, fna, fe, fg;
  bool ok = slurp(base + "graph_indicator.txt", &fi, err) && slurp(base + "A.txt", &fa, err) &&
            slurp(base + "node_labels.txt", &fn, err) && slurp(base + "edge_labels.txt", &fe, err) &&
            slurp(base + "graph_labels.txt", &fg, err);
  if (ok && (flags & GK_TU_ATTR_NODES)) ok = slurp(base + "node_attributes.txt", &fna, err);
  if (ok && !fi.present) { *err = "missing " + base + "graph_indicator.txt"; ok = false; }
  if (ok && !fa.present) { *err = "missing " + base + "A.txt"; ok = false; }
  std::vector<int32_t> dummy;
  if (ok) ok = parse_ints(fi, 1, &t->indicator, &dummy, "graph_indicator", err);
  if (ok) ok = parse_ints(fa, 2, &t->eu, &t->ev, "A", err);
  // read_data only opens the label file when it does not take the attributes (base.py:223-240)
  if (ok && fn.present && !fna.present) ok = parse_ints(fn, 1, &t->node_label, &dummy, "node_labels", err);
  if (ok && fe.present) ok = parse_ints(fe, 1, &t->el, &dummy, "edge_labels", err);
  if (ok && fg.present) ok = parse_ints(fg, 1, &t->classes, &dummy, "graph_labels", err);
  if (ok && fna.present) ok = parse_floats(fna, &t->node_attr, &t->attr_dim, "node_attributes", err);
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

struct DirEdge {
  int32_t u, v, lab;
  uint32_t seq;  // write order in the reference: line e forward = 2e, reverse = 2e + 1
};

// mode GK_TU_LABELLED_NODES: vertex set = the keys of the node-label dictionary (what WeisfeilerLehman walks,
//   weisfeiler_lehman.py:234) -- every node of the graph when a label / attribute file exists;
// mode GK_TU_EDGE_NODES: vertex set = the nodes that occur in an edge of their graph (ShortestPath's
//   sorted edge symbols, graph.py:1613-1631; the edge-dictionary keys WL-OA walks).
inline int pack(gk_tu* t, int32_t mode, std::string* err) {
  if (mode != GK_TU_LABELLED_NODES && mode != GK_TU_EDGE_NODES) { *err = "gk_tu_pack: unknown mode"; return GK_ERR_ARG; }
  const int64_t n = t->n_nodes;
  const bool sym = t->flags & GK_TU_SYMMETRIC;
  const bool has_el = !t->el.empty();
  std::vector<DirEdge> de;
  de.reserve(t->eu.size() * (sym ? 2 : 1));
  for (size_t e = 0; e < t->eu.size(); ++e) {
    const int32_t u = t->eu[e], v = t->ev[e];
    if (t->indicator[u] != t->indicator[v]) {
      *err = "A: line " + std::to_string(e + 1) + " joins nodes of different graphs (the reference raises KeyError on it)";
      return GK_ERR_ARG;
    }
    const int32_t lab = has_el ? t->el[e] : 0;
    de.push_back({u, v, lab, (uint32_t)(2 * e)});
    if (sym) de.push_back({v, u, lab, (uint32_t)(2 * e + 1)});
  }
  {  // sort by (u, v, write order): counting sort on the source, then each (short) row on its own
    std::vector<uint32_t> start(n + 1, 0);
    for (const auto& e : de) start[e.u + 1] += 1;
    for (int64_t i = 0; i < n; ++i) start[i + 1] += start[i];
    std::vector<DirEdge> tmp(de.size());
    std::vector<uint32_t> cur(start.begin(), start.end() - 1);
    for (const auto& e : de) tmp[cur[e.u]++] = e;
    de.swap(tmp);
    for (int64_t i = 0; i < n; ++i) {
      const uint32_t lo = start[i], hi = start[i + 1];
      if (hi - lo > 32) {
        std::sort(de.begin() + lo, de.begin() + hi, [](const DirEdge& a, const DirEdge& b) {
          if (a.v != b.v) return a.v < b.v;
          return a.seq < b.seq;
        });
      } else {
        for (uint32_t x = lo + 1; x < hi; ++x) {  // insertion sort: rows are a handful of edges
          const DirEdge e = de[x];
          uint32_t y = x;
          while (y > lo && (de[y - 1].v > e.v || (de[y - 1].v == e.v && de[y - 1].seq > e.seq))) { de[y] = de[y - 1]; --y; }
          de[y] = e;
        }
      }
    }
  }
  size_t w = 0;  // keep the LAST write of every (u, v)
  for (size_t i = 0; i < de.size(); ++i) {
    if (i + 1 < de.size() && de[i + 1].u == de[i].u && de[i + 1].v == de[i].v) continue;
    de[w++] = de[i];
  }
  de.resize(w);
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
    for (const auto& e : de)
      if (e.u != e.v) { label[e.u] += 1; labelled[e.u] = 1; }
  }
  std::vector<char> keep(n, 0);
  if (mode == GK_TU_LABELLED_NODES) {
    keep = labelled;
  } else {
    for (const auto& e : de) { keep[e.u] = 1; keep[e.v] = 1; }
  }
  // vertices grouped by graph, node order inside a graph (stable counting sort)
  t->graph_ptr.assign(t->n_graphs + 1, 0);
  for (int64_t i = 0; i < n; ++i)
    if (keep[i]) t->graph_ptr[t->indicator[i] + 1] += 1;
  for (int64_t g = 0; g < t->n_graphs; ++g) t->graph_ptr[g + 1] += t->graph_ptr[g];
  const int64_t V = t->graph_ptr[t->n_graphs];
  std::vector<int32_t> pos(n, -1), cursor(t->graph_ptr.begin(), t->graph_ptr.end() - 1);
  t->vnode.assign(V, 0);
  for (int64_t i = 0; i < n; ++i)
    if (keep[i]) { pos[i] = cursor[t->indicator[i]]++; t->vnode[pos[i]] = (int32_t)i; }
  // CSR over the kept vertices.  An edge whose source is not a vertex is never visited by the reference
  // (weisfeiler_lehman.py:234 walks labelled vertices only); a target that is not a vertex is its KeyError.
  t->row_ptr.assign(V + 1, 0);
  size_t kept_edges = 0;
  for (const auto& e : de) {
    if (pos[e.u] < 0) continue;
    if (pos[e.v] < 0) {
      *err = "node " + std::to_string(e.v + 1) + " is the target of an edge but has no label (the reference raises KeyError)";
      return GK_ERR_ARG;
    }
    t->row_ptr[pos[e.u] + 1] += 1;
    ++kept_edges;
  }
  for (int64_t v = 0; v < V; ++v) t->row_ptr[v + 1] += t->row_ptr[v];
  t->col_idx.assign(kept_edges, 0);
  t->elabel.assign(has_el ? kept_edges : 0, 0);
  std::vector<int32_t> fill(t->row_ptr.begin(), t->row_ptr.end() - 1);
  // `de` is sorted by (u, v) and positions are monotone in the node id inside a graph, so every row comes out sorted
  for (const auto& e : de) {
    if (pos[e.u] < 0) continue;
    const int32_t k = fill[pos[e.u]]++;
    t->col_idx[k] = pos[e.v];
    if (has_el) t->elabel[k] = e.lab;
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
