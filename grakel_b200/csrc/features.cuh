// Sparse feature block -> column statistics -> dense bf16 panel.
//
// The feature block is the (graph, column) -> count hash table filled by the WL /
// SP feature kernels: it is the reference's per-level `csr_matrix` Phi_i
// (vertex_histogram.py:125-150) for all levels side by side, or the dense
// `phi_x` of ShortestPath (shortest_path.py:396-400).
//
// Only columns that can contribute to an off-diagonal entry are contracted on the
// tensor cores:
//    square case  (fit_transform): columns present in >= 2 graphs;
//    rectangular  (transform)    : columns present in an X graph AND a Y graph
// (this is exactly the `Y[:, :X.shape[1]]` slice of vertex_histogram.py:179 --
// unseen columns are dropped).  Self similarities use ALL columns
// (vertex_histogram.py:186-219) and are computed here as exact integers.
#pragma once
#include "common.cuh"
#include "wl.cuh"

namespace gk {

// pass 1 over the table: per-column first/last graph, per-graph sum of squares,
// global nnz / max count.
__global__ void __launch_bounds__(256)
feat_pass1(size_t cap, const unsigned long long* __restrict__ keys, const unsigned* __restrict__ cnt,
           int* colfirst, int* collast, unsigned long long* diag, DevScalars* sc) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned my = 0;
  unsigned has = 0;
  if (i < cap) {
    unsigned long long k = keys[i];
    if (k != EMPTY64) {
      const int g = (int)(k >> 32);
      const unsigned c = (unsigned)k;
      my = cnt[i];
      has = 1;
      atomicMin(&colfirst[c], g);
      atomicMax(&collast[c], g);
      atomicAdd(&diag[g], (unsigned long long)my * my);
    }
  }
  // block-level reductions before the global atomics
  unsigned mx = my, n = has;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    n += __shfl_xor_sync(0xffffffffu, n, d);
  }
  if ((threadIdx.x & 31) == 0 && n) {
    atomicMax(&sc->max_count, (unsigned long long)mx);
    atomicAdd(&sc->n_entries, (unsigned long long)n);
  }
}

__global__ void __launch_bounds__(256)
diag_finish(int N, const unsigned long long* __restrict__ diag, double* __restrict__ diag_f64,
            DevScalars* sc) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long d = 0;
  if (g < N) {
    d = diag[g];
    diag_f64[g] = (double)d;
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    unsigned long long o = __shfl_xor_sync(0xffffffffu, d, s);
    d = d > o ? d : o;
  }
  if ((threadIdx.x & 31) == 0 && d) atomicMax(&sc->max_diag, d);
}

// dense-column selection flags + per-block sums (first half of the column scan)
__global__ void __launch_bounds__(256)
col_flags(long long D, int n_fit, int N, const int* __restrict__ colfirst,
          const int* __restrict__ collast, int* __restrict__ flags, int* __restrict__ block_sums) {
  long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int f = 0;
  if (c < D) {
    const int a = colfirst[c], b = collast[c];
    if (b >= 0) f = (n_fit >= N) ? (a != b) : (a < n_fit && b >= n_fit);
    flags[c] = f;
  }
  int total;
  block_exclusive_scan_256(f, &total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(256)
col_assign(long long D, const int* __restrict__ flags, const int* __restrict__ block_sums,
           int* __restrict__ dense_col) {
  long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int f = c < D ? flags[c] : 0;
  int total;
  const int ex = block_exclusive_scan_256(f, &total);
  if (c < D) dense_col[c] = f ? block_sums[blockIdx.x] + ex : -1;
}

// pass 2 over the table: scatter the selected columns into the zeroed bf16 panel
__global__ void __launch_bounds__(256)
feat_fill_panel(size_t cap, const unsigned long long* __restrict__ keys, const unsigned* __restrict__ cnt,
                const int* __restrict__ dense_col, __nv_bfloat16* __restrict__ panel, long long ld) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cap) return;
  unsigned long long k = keys[i];
  if (k == EMPTY64) return;
  const int dc = dense_col[(unsigned)k];
  if (dc < 0) return;
  panel[(long long)(k >> 32) * ld + dc] = __float2bfloat16_rn((float)cnt[i]);
}

// same, into a u32 panel (exact CUDA-core Gram)
__global__ void __launch_bounds__(256)
feat_fill_panel_u32(size_t cap, const unsigned long long* __restrict__ keys, const unsigned* __restrict__ cnt,
                    const int* __restrict__ dense_col, unsigned* __restrict__ panel, long long ld) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cap) return;
  unsigned long long k = keys[i];
  if (k == EMPTY64) return;
  const int dc = dense_col[(unsigned)k];
  if (dc < 0) return;
  panel[(long long)(k >> 32) * ld + dc] = cnt[i];
}

}  // namespace gk
