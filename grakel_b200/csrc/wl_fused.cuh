// All WL levels in ONE persistent cooperative kernel (weisfeiler_lehman.py:199-258 +
// vertex_histogram.py:107-122 for every level), staged through shared memory.
//
// Why: the multi-kernel pipeline of wl.cuh costs ~80 us per level in five latency-bound launches
// (profiles/r01b_full_summary.md) against an HBM floor of 2 us.  Micro-benchmarks
// (tools/mb/wl_mb.cu) show the limit is the number of SCATTERED L2 operations per vertex
// (~150-300 G sector ops/s chip-wide), so this kernel is organised to minimise them:
//
//   * The graphs are cut into TILES of whole graphs (<= WLF_TILE_V vertices, <= WLF_TILE_E edges);
//     each CTA owns a contiguous run of tiles.  A tile's CSR slice and labels are staged into
//     shared memory with coalesced loads; neighbour gathers, sorting and the per-graph label
//     histogram never leave shared memory.
//   * Signature dedup uses ONE packed 64-bit word per slot, {32-bit tag | representative vertex}:
//     insert = 1 L2 read + (1 CAS when the slot is empty | 1 atomicMin when the tag matches and the
//     current representative is larger | nothing).  The full signature is still verified against
//     the representative's, so the partition is exact; a tag clash only triggers the host's retry.
//   * (graph, column) counts are produced by comparing labels inside the graph in shared memory
//     (a graph never straddles a tile) and appended to a compact COO list -- no global feature
//     hash table, no per-level table clearing, and the Gram stage streams 12 B per entry instead
//     of 12 B per (4x over-provisioned) slot.
//
// A level is   [A] signatures + insert | grid barrier | [B] representative + verification + ranks
//              | grid barrier | [C] dense first-occurrence ids, labels, feature entries
// and [A] of the next level follows [C] after a __syncthreads only, because everything a tile
// needs from the previous level (labels of its own vertices) was written by the same CTA.
#pragma once
#include "common.cuh"
#include "wl.cuh"

namespace gk {

constexpr int WLF_THREADS = 1024;
constexpr int WLF_TILE_V = 4096;    // vertices per tile (u16 local indices)
constexpr int WLF_TILE_E = 16384;   // edges per tile
constexpr int WLF_RANK_BITS = 20;   // rank_pack = cta << 20 | rank of the representative inside its CTA
constexpr int WLF_SMEM = (WLF_TILE_V + 1) * 4 /*rp_s*/ + WLF_TILE_E * 2 /*col_s*/ + WLF_TILE_V * 4 /*lab_s*/ +
                         WLF_TILE_E * 4 /*sig_s*/ + 16;

struct WlFusedParams {
  int V, L;
  const int* graph_ptr;
  const int* row_ptr;
  const int* col_idx;
  const int* vgraph;
  const int* labels0;
  const int* tile_vbeg;  // [n_tiles + 1] first vertex of each tile (whole graphs)
  const int* cta_tile;   // [grid + 1] first tile of each CTA
  int* labels_all;       // [L * V]
  int* sig_nbr;          // [E] sorted neighbour labels of the current level
  int* slot_of;          // [V] hash slot, then representative vertex
  int* rank_pack;        // [V] valid for representatives
  int* cta_count;        // [grid]
  unsigned long long* table;  // 2 x (ht_mask + 1) packed {tag, representative} words
  unsigned ht_mask;
  unsigned long long* coo_keys;  // feature entries: graph << 32 | column
  unsigned* coo_cnt;
  unsigned* barrier;  // zeroed by the host before the launch
  unsigned long long seed;
  FeatStats st;
  DevScalars* sc;   // sc->sp_coo is the COO append counter
  long long* prof;  // optional [grid][L][16] globaltimer stamps (GRAKEL_B200_PROF), else NULL
};

__device__ __forceinline__ long long wlf_now() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define WLF_STAMP(level, k) \
  do { if (p.prof && tid == 0) p.prof[((size_t)b * p.L + (level)) * 16 + (k)] = wlf_now(); } while (0)

__device__ __forceinline__ void wlf_grid_sync(unsigned* bar, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    unsigned seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar) : "memory");
    } while (seen < target);
    __threadfence();
  }
  __syncthreads();
}

// block-wide exclusive scan for 1024 threads; returns the exclusive prefix, *total = block sum
__device__ __forceinline__ int wlf_block_scan(int x, int* total, int* s_warp /*[32]*/) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int incl = x;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += y;
  }
  if (lane == 31) s_warp[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    int s = s_warp[lane];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, s, d);
      if (lane >= d) s += y;
    }
    s_warp[lane] = s;
  }
  __syncthreads();
  const int off = wid ? s_warp[wid - 1] : 0;
  *total = s_warp[31];
  __syncthreads();
  return off + incl - x;
}

// packed-table insert: returns the slot whose word carries this signature's tag
__device__ __forceinline__ unsigned wlf_insert(unsigned long long* tab, unsigned mask, unsigned long long key, int v) {
  const unsigned long long mine = (key & 0xFFFFFFFF00000000ULL) | (unsigned)v;
  unsigned slot = (unsigned)((key & 0xFFFFFFFFULL) * 0x9E3779B1ULL >> 8) & mask;
  while (true) {
    unsigned long long w = __ldcg(&tab[slot]);
    if (w == EMPTY64) {
      w = atomicCAS(&tab[slot], EMPTY64, mine);
      if (w == EMPTY64) return slot;
    }
    if ((w >> 32) == (key >> 32)) {
      if ((unsigned)w > (unsigned)v) atomicMin(&tab[slot], mine);
      return slot;
    }
    slot = (slot + 1) & mask;
  }
}

// stage one tile: CSR slice (row offsets relative to the tile, neighbours as tile-local u16)
__device__ __forceinline__ void wlf_stage(const WlFusedParams& p, int v0, int nv, int e0, int ne, const int* lab_src,
                                          int* rp_s, unsigned short* col_s, int* lab_s, bool with_csr) {
  const int tid = threadIdx.x;
  if (with_csr) {
    for (int i = tid; i <= nv; i += WLF_THREADS) rp_s[i] = p.row_ptr[v0 + i] - e0;
    for (int k = tid; k < ne; k += WLF_THREADS) col_s[k] = (unsigned short)(p.col_idx[e0 + k] - v0);
  }
  if (lab_src)
    for (int i = tid; i < nv; i += WLF_THREADS) lab_s[i] = lab_src[v0 + i];
}

// feature entries of one tile from the labels in lab_s: vertex i emits (graph, base + label, count)
// iff it is the first vertex of its graph carrying that label.
__device__ __forceinline__ void wlf_emit(const WlFusedParams& p, int v0, int nv, const int* lab_s, long long base,
                                         int* s_warp, int* s_bcast, unsigned& mx, unsigned& n_new) {
  const int tid = threadIdx.x, lane = tid & 31;
  for (int i0 = 0; i0 < nv; i0 += WLF_THREADS) {
    const int i = i0 + tid;
    int emit = 0, g = -1;
    unsigned cnt = 0;
    int l = 0;
    if (i < nv) {
      g = p.vgraph[v0 + i];
      const int gs = p.graph_ptr[g] - v0, ge = p.graph_ptr[g + 1] - v0;
      l = lab_s[i];
      bool first = true;
      for (int u = gs; u < ge; ++u) {
        const bool same = lab_s[u] == l;
        cnt += same ? 1u : 0u;
        first = first && !(same && u < i);
      }
      emit = first ? 1 : 0;
    }
    int total;
    const int ex = wlf_block_scan(emit, &total, s_warp);
    if (tid == 0) *s_bcast = total ? (int)atomicAdd(&p.sc->sp_coo, (unsigned long long)total) : 0;
    __syncthreads();
    const int off = *s_bcast;
    if (emit) {
      const unsigned long long col = (unsigned long long)(base + l);
      p.coo_keys[off + ex] = ((unsigned long long)(unsigned)g << 32) | col;
      p.coo_cnt[off + ex] = cnt;
      if (__ldcg(&p.st.colcnt[col]) < COL_CAP) atomicAdd(&p.st.colcnt[col], 1u);
      mx = max(mx, cnt);
      n_new += 1u;
    }
    // exact self similarity: sum of squared counts per graph (runs of equal g inside the warp)
    unsigned long long val = emit ? (unsigned long long)cnt * cnt : 0ULL;
    const int gg = g;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned long long y = __shfl_down_sync(0xffffffffu, val, d);
      const int gy = __shfl_down_sync(0xffffffffu, gg, d);
      if (lane + d < 32 && gy == gg) val += y;
    }
    const int gprev = __shfl_up_sync(0xffffffffu, gg, 1);
    if (g >= 0 && (lane == 0 || gprev != gg) && val) atomicAdd(&p.st.diag[g], val);
    __syncthreads();  // s_bcast is reused by the next round
  }
}

__global__ void __launch_bounds__(WLF_THREADS, 1)
wl_fused_kernel(WlFusedParams p) {
  extern __shared__ __align__(16) unsigned char wlf_smem[];
  int* rp_s = reinterpret_cast<int*>(wlf_smem);
  int* lab_s = rp_s + (WLF_TILE_V + 1);
  int* sig_s = lab_s + WLF_TILE_V;
  unsigned short* col_s = reinterpret_cast<unsigned short*>(sig_s + WLF_TILE_E);
  __shared__ int s_warp[32];
  __shared__ int s_prefix[1024];  // exclusive scan of the CTA counts (grid <= 1024)
  __shared__ unsigned s_red[64];
  __shared__ int s_bcast;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int G = gridDim.x, b = blockIdx.x;
  const int t_beg = p.cta_tile[b], t_end = p.cta_tile[b + 1];
  const int V = p.V;
  const size_t ht_cap = (size_t)p.ht_mask + 1;
  unsigned n_sync = 0;
  long long level_base = 0;  // first column of the current level (uniform across the grid)
  unsigned mx = 0, n_new = 0;
  WLF_STAMP(0, 0);

  // publish the per-CTA partials of a level (max count, created entries)
  auto flush_partials = [&](int level) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
      n_new += __shfl_xor_sync(0xffffffffu, n_new, d);
    }
    if (lane == 0) { s_red[wid] = mx; s_red[32 + wid] = n_new; }
    __syncthreads();
    if (tid == 0) {
      unsigned m = 0, n = 0;
      for (int w = 0; w < 32; ++w) { m = max(m, s_red[w]); n += s_red[32 + w]; }
      p.st.part_max[(size_t)level * G + b] = m;
      p.st.part_new[(size_t)level * G + b] = n;
    }
    mx = 0;
    n_new = 0;
    __syncthreads();
  };

  // ---- level 0: labels as given
  for (int t = t_beg; t < t_end; ++t) {
    const int v0 = p.tile_vbeg[t], nv = p.tile_vbeg[t + 1] - v0;
    wlf_stage(p, v0, nv, 0, 0, p.labels0, rp_s, col_s, lab_s, false);
    for (int i = tid; i < nv; i += WLF_THREADS) p.labels_all[v0 + i] = p.labels0[v0 + i];
    __syncthreads();
    wlf_emit(p, v0, nv, lab_s, 0, s_warp, &s_bcast, mx, n_new);
  }
  level_base = p.sc->level_base[1];  // = number of level-0 labels (set by the host)
  flush_partials(0);
  WLF_STAMP(0, 6);

  for (int lv = 1; lv < p.L; ++lv) {
    const int* lab_in = p.labels_all + (size_t)(lv - 1) * V;
    int* lab_out = p.labels_all + (size_t)lv * V;
    unsigned long long* tab = p.table + (size_t)(lv & 1) * ht_cap;

    // ---------------- [A] signatures + insert
    WLF_STAMP(lv, 0);
    for (int t = t_beg; t < t_end; ++t) {
      const int v0 = p.tile_vbeg[t], nv = p.tile_vbeg[t + 1] - v0;
      const int e0 = p.row_ptr[v0], ne = p.row_ptr[v0 + nv] - e0;
      __syncthreads();  // previous tile's shared memory is no longer read
      wlf_stage(p, v0, nv, e0, ne, lab_in, rp_s, col_s, lab_s, true);
      __syncthreads();
      WLF_STAMP(lv, 8);
      // one thread per vertex of degree <= 8: 19-comparator network in registers
      for (int i = tid; i < nv; i += WLF_THREADS) {
        const int beg = rp_s[i];
        const int deg = rp_s[i + 1] - beg;
        if (deg > 8) continue;
        int x0, x1, x2, x3, x4, x5, x6, x7;
        x0 = 0 < deg ? lab_s[col_s[beg + 0]] : 0x7fffffff;
        x1 = 1 < deg ? lab_s[col_s[beg + 1]] : 0x7fffffff;
        x2 = 2 < deg ? lab_s[col_s[beg + 2]] : 0x7fffffff;
        x3 = 3 < deg ? lab_s[col_s[beg + 3]] : 0x7fffffff;
        x4 = 4 < deg ? lab_s[col_s[beg + 4]] : 0x7fffffff;
        x5 = 5 < deg ? lab_s[col_s[beg + 5]] : 0x7fffffff;
        x6 = 6 < deg ? lab_s[col_s[beg + 6]] : 0x7fffffff;
        x7 = 7 < deg ? lab_s[col_s[beg + 7]] : 0x7fffffff;
        GK_CSWAP(x0, x1) GK_CSWAP(x2, x3) GK_CSWAP(x4, x5) GK_CSWAP(x6, x7)
        GK_CSWAP(x0, x2) GK_CSWAP(x1, x3) GK_CSWAP(x4, x6) GK_CSWAP(x5, x7)
        GK_CSWAP(x1, x2) GK_CSWAP(x5, x6) GK_CSWAP(x0, x4) GK_CSWAP(x3, x7)
        GK_CSWAP(x1, x5) GK_CSWAP(x2, x6)
        GK_CSWAP(x1, x4) GK_CSWAP(x3, x6)
        GK_CSWAP(x2, x4) GK_CSWAP(x3, x5)
        GK_CSWAP(x3, x4)
        const int xs[8] = {x0, x1, x2, x3, x4, x5, x6, x7};
        unsigned long long tt = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (j < deg) {
            sig_s[beg + j] = xs[j];
            tt += sig_term(xs[j], j, p.seed);
          }
        }
        p.slot_of[v0 + i] = (int)wlf_insert(tab, p.ht_mask, sig_final(tt, lab_s[i], deg, p.seed), v0 + i);
      }
      if (p.prof) { __syncthreads(); WLF_STAMP(lv, 9); }
      // higher degrees: one warp per vertex, all-ascending bitonic network in the vertex's own
      // shared-memory segment (valid for any length: exchanges with the virtual +inf tail are no-ops)
      for (int i0 = wid * 32; i0 < nv; i0 += WLF_THREADS) {
        const int iv = i0 + lane;
        const bool big = iv < nv && (rp_s[iv + 1] - rp_s[iv]) > 8;
        unsigned m = __ballot_sync(0xffffffffu, big);
        while (m) {
          const int i = i0 + __ffs(m) - 1;
          m &= m - 1;
          const int beg = rp_s[i], deg = rp_s[i + 1] - beg;
          int* seg = sig_s + beg;
          for (int j = lane; j < deg; j += 32) seg[j] = lab_s[col_s[beg + j]];
          __syncwarp();
          int n2 = 1;
          while (n2 < deg) n2 <<= 1;
          for (int k = 2; k <= n2; k <<= 1) {
            for (int j = lane; j < deg; j += 32) {
              const int q = j ^ (k - 1);
              if (q > j && q < deg) {
                const int a = seg[j], c = seg[q];
                if (a > c) { seg[j] = c; seg[q] = a; }
              }
            }
            __syncwarp();
            for (int s = k >> 2; s > 0; s >>= 1) {
              for (int j = lane; j < deg; j += 32) {
                const int q = j ^ s;
                if (q > j && q < deg) {
                  const int a = seg[j], c = seg[q];
                  if (a > c) { seg[j] = c; seg[q] = a; }
                }
              }
              __syncwarp();
            }
          }
          unsigned long long tt = 0;
          for (int j = lane; j < deg; j += 32) tt += sig_term(seg[j], j, p.seed);
#pragma unroll
          for (int s = 16; s > 0; s >>= 1) tt += __shfl_xor_sync(0xffffffffu, tt, s);
          if (lane == 0)
            p.slot_of[v0 + i] = (int)wlf_insert(tab, p.ht_mask, sig_final(tt, lab_s[i], deg, p.seed), v0 + i);
        }
      }
      __syncthreads();
      WLF_STAMP(lv, 10);
      // sorted neighbour labels to global memory (coalesced): other CTAs verify against them
      for (int k = tid; k < ne; k += WLF_THREADS) p.sig_nbr[e0 + k] = sig_s[k];
    }
    __syncthreads();
    WLF_STAMP(lv, 1);
    wlf_grid_sync(p.barrier, (++n_sync) * (unsigned)G);
    WLF_STAMP(lv, 2);

    // ---------------- [B] representative, verification, CTA-local rank
    int carry = 0;
    for (int t = t_beg; t < t_end; ++t) {
      const int v0 = p.tile_vbeg[t], nv = p.tile_vbeg[t + 1] - v0;
      for (int i0 = 0; i0 < nv; i0 += WLF_THREADS) {
        const int v = v0 + i0 + tid;
        int f = 0;
        if (i0 + tid < nv) {
          const int r = (int)(unsigned)__ldcg(&tab[p.slot_of[v]]);
          p.slot_of[v] = r;
          f = (r == v);
          if (!f) {
            const int bv = p.row_ptr[v], dv = p.row_ptr[v + 1] - bv;
            const int br = p.row_ptr[r], dr = p.row_ptr[r + 1] - br;
            bool same = (dv == dr) && (lab_in[v] == __ldcg(&lab_in[r]));
            for (int j = 0; same && j < dv; ++j) same = p.sig_nbr[bv + j] == __ldcg(&p.sig_nbr[br + j]);
            if (!same) atomicOr(&p.sc->collision, 1u);
          }
        }
        int total;
        const int ex = wlf_block_scan(f, &total, s_warp);
        if (f) p.rank_pack[v] = (b << WLF_RANK_BITS) | (carry + ex);
        carry += total;
      }
    }
    if (tid == 0) p.cta_count[b] = carry;
    __syncthreads();
    WLF_STAMP(lv, 3);
    if (lv + 1 < p.L) {  // clear the table the next level inserts into (last read in [B] of level lv-1)
      const uint4 ones = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
      uint4* hk = reinterpret_cast<uint4*>(p.table + (size_t)((lv + 1) & 1) * ht_cap);
      for (size_t i = (size_t)b * WLF_THREADS + tid; i < ht_cap / 2; i += (size_t)G * WLF_THREADS) hk[i] = ones;
    }
    __syncthreads();
    WLF_STAMP(lv, 4);
    wlf_grid_sync(p.barrier, (++n_sync) * (unsigned)G);
    WLF_STAMP(lv, 5);

    // ---------------- [C] dense ids, labels, feature entries
    {
      const int c = tid < G ? __ldcg(&p.cta_count[tid]) : 0;
      int total;
      const int ex = wlf_block_scan(c, &total, s_warp);
      s_prefix[tid] = ex;
      __syncthreads();
      for (int t = t_beg; t < t_end; ++t) {
        const int v0 = p.tile_vbeg[t], nv = p.tile_vbeg[t + 1] - v0;
        for (int i = tid; i < nv; i += WLF_THREADS) {
          const int v = v0 + i;
          const int r = p.slot_of[v];
          const int rp = __ldcg(&p.rank_pack[r]);
          const int id = s_prefix[rp >> WLF_RANK_BITS] + (rp & ((1 << WLF_RANK_BITS) - 1));
          lab_out[v] = id;
          lab_s[i] = id;
        }
        __syncthreads();
        wlf_emit(p, v0, nv, lab_s, level_base, s_warp, &s_bcast, mx, n_new);
      }
      if (b == 0 && tid == 0) {
        p.sc->level_dims[lv] = total;
        p.sc->level_base[lv + 1] = level_base + total;
      }
      level_base += total;
      flush_partials(lv);  // ends with __syncthreads: labels of this level are complete for the CTA's graphs
      WLF_STAMP(lv, 6);
    }
  }
}

}  // namespace gk
