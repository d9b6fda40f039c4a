// Weisfeiler-Lehman optimal assignment (weisfeiler_lehman_optimal_assignment.py:78-279) as a Gram.
//
// The reference builds, per graph, the vector Hs[j, c] = number of vertices of j that carry WL label c
// at c's level (the hierarchy walk of :203-209 adds 1 to every ancestor of a vertex's last-level
// label), i.e. exactly the WL feature block of all levels, and then fills K with an O(N^2) Python loop
// of histogram intersections  K[i, j] = sum_c min(Hs[i, c], Hs[j, c])  (:257-266).
//
// min(a, b) = sum_{t >= 1} [a >= t][b >= t]  for non-negative integers, so the intersection kernel is
// the DOT PRODUCT of the unary ("thermometer") expansions: column c with count k becomes the k columns
// (c, 1) ... (c, k), each holding 1.  The expanded block has exactly one entry per (vertex, level) --
// (h + 1) V entries, barely more than the nnz of the count block -- and feeds the same head / tail Gram
// as the subtree kernel (features.cuh, gram_tc.cuh): high-frequency threshold columns go through the
// tcgen05 GEMM, the long tail through exact pair updates, self similarities are sums of 1 = (h + 1) n_j,
// and the transform rule "drop the columns X does not have" (:433) is the rectangular column filter.
// Everything stays integer-exact.
#pragma once
#include "common.cuh"
#include "wl.cuh"

namespace gk {

struct OaCursors {
  unsigned long long n_cols;     // columns of the expanded block
  unsigned long long n_entries;  // entries of the expanded block
};

// pass 1: largest count of every column (L2-read filter: most updates do not raise the maximum)
__global__ void __launch_bounds__(256)
oa_colmax(size_t cap, const unsigned long long* __restrict__ keys, const unsigned* __restrict__ cnt, unsigned* colmax) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += (size_t)gridDim.x * blockDim.x) {
    const unsigned long long k = keys[i];
    if (k == EMPTY64) continue;
    const unsigned c = (unsigned)k, n = cnt[i];
    if (__ldcg(&colmax[c]) < n) atomicMax(&colmax[c], n);
  }
}

constexpr int OA_THREADS = 1024;

// Block-wide reservation from one global cursor: exclusive offset of this thread's `n` items.  ONE atomic per
// block -- returning atomics on a single address serialise in their L2 slice (a warp-aggregated version of
// oa_expand spent ~200 us in 74 k of them, profiles/r01h_full_summary.md).  All threads of the block must call.
__device__ __forceinline__ unsigned long long oa_block_reserve(unsigned n, unsigned long long* cursor) {
  __shared__ unsigned s_tot[32];
  __shared__ unsigned long long s_base;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  unsigned incl = n;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned y = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += y;
  }
  if (lane == 31) s_tot[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    const unsigned t = lane < (int)(blockDim.x >> 5) ? s_tot[lane] : 0u;
    unsigned i2 = t;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned y = __shfl_up_sync(0xffffffffu, i2, d);
      if (lane >= d) i2 += y;
    }
    s_tot[lane] = i2 - t;  // exclusive offset of warp `lane`
    if (lane == 31) s_base = i2 ? atomicAdd(cursor, (unsigned long long)i2) : 0ULL;
  }
  __syncthreads();
  return s_base + s_tot[wid] + incl - n;
}

// pass 2: first threshold column of every column.  The order of the expanded columns depends on
// scheduling, K does not (exact integer arithmetic).
__global__ void __launch_bounds__(OA_THREADS)
oa_colbase(long long D, const unsigned* __restrict__ colmax, unsigned* __restrict__ colbase, OaCursors* cur) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned m = c < D ? colmax[c] : 0u;
  const unsigned long long base = oa_block_reserve(m, &cur->n_cols);
  if (c < D) colbase[c] = (unsigned)base;
}

// pass 3: entry (graph, c, k) -> k entries (graph, colbase[c] + t, 1), t < k; statistics of the new
// block (graphs per column, self similarity = number of entries of the graph) maintained on the fly.
// Graphs-per-column counts are aggregated in a shared-memory table first: a threshold column of a
// level-0 label is held by thousands of graphs, and same-address global atomics retire one after the
// other at ~28 ns each (the un-aggregated version spent 230 us in chains of up to 8 192 of them,
// profiles/r01i_launches_wloa.csv); per block a hot column now costs one global atomic.
constexpr int OA_AGG = 4096;  // slots of the per-block (column -> graphs) table

__global__ void __launch_bounds__(OA_THREADS)
oa_expand(size_t cap, const unsigned long long* __restrict__ keys, const unsigned* __restrict__ cnt,
          const unsigned* __restrict__ colbase, size_t out_cap, unsigned long long* __restrict__ out_keys,
          unsigned* __restrict__ out_cnt, OaCursors* cur, unsigned* colcnt, unsigned long long* diag, DevScalars* sc) {
  __shared__ unsigned agg_col[OA_AGG], agg_cnt[OA_AGG];
  for (int s = threadIdx.x; s < OA_AGG; s += OA_THREADS) { agg_col[s] = 0xFFFFFFFFu; agg_cnt[s] = 0u; }
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  unsigned long long k = EMPTY64;
  unsigned n = 0;
  if (i < cap) {
    k = keys[i];
    if (k != EMPTY64) n = cnt[i];
  }
  const unsigned long long pos = oa_block_reserve(n, &cur->n_entries);  // its barriers also order the table init
  // self similarity: one atomic per (warp, graph) -- COO entries of a graph are adjacent
  const unsigned g = n ? (unsigned)(k >> 32) : 0xffffffffu;
  const unsigned peers = __match_any_sync(0xffffffffu, g);
  unsigned gsum = 0;
  for (unsigned m = peers; m; m &= m - 1) gsum += __shfl_sync(peers, n, __ffs(m) - 1);
  if (n && lane == __ffs(peers) - 1) atomicAdd(&diag[g], (unsigned long long)gsum);
  if (n) {
    if (pos + n > out_cap) {
      sc->ft_overflow = 1u;
    } else {
      const unsigned b = colbase[(unsigned)k];
      const unsigned long long hi = k & 0xffffffff00000000ULL;
      for (unsigned t = 0; t < n; ++t) {
        const unsigned col = b + t;
        out_keys[pos + t] = hi | (unsigned long long)col;
        out_cnt[pos + t] = 1u;
        unsigned slot = (col * 0x9E3779B1u >> 14) & (OA_AGG - 1);
        bool done = false;
        for (int probe = 0; probe < 16 && !done; ++probe) {
          unsigned prev = agg_col[slot];
          if (prev == 0xFFFFFFFFu) prev = atomicCAS(&agg_col[slot], 0xFFFFFFFFu, col);
          if (prev == 0xFFFFFFFFu || prev == col) { atomicAdd(&agg_cnt[slot], 1u); done = true; }
          else slot = (slot + 1) & (OA_AGG - 1);
        }
        if (!done && __ldcg(&colcnt[col]) < COL_CAP) atomicAdd(&colcnt[col], 1u);  // table crowded: straight to global
      }
    }
  }
  __syncthreads();
  for (int s = threadIdx.x; s < OA_AGG; s += OA_THREADS) {
    const unsigned col = agg_col[s];
    if (col != 0xFFFFFFFFu && __ldcg(&colcnt[col]) < COL_CAP) atomicAdd(&colcnt[col], agg_cnt[s]);
  }
}

// per-CTA partials of the feature kernels, as diag_finish folds them: one "CTA", largest count 1
__global__ void oa_finish(const OaCursors* cur, unsigned* part_max, unsigned* part_new, const DevScalars* sc, int add_frozen) {
  // frozen vertices of wl_fused2 wrote no count entries; each of their (vertex, level) units is one unary entry
  const unsigned long long n = cur->n_entries + (add_frozen ? sc->sp_coo : 0ULL);
  part_max[0] = n ? 1u : 0u;
  part_new[0] = (unsigned)n;
}

}  // namespace gk
