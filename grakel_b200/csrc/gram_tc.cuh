// Gram matrix K = A * B^T over the dense bf16 feature panel on the 5th-gen tensor
// cores (tcgen05.mma, fp32 accumulators in TMEM, operands staged by TMA with the
// 128-byte swizzle).  Replaces `self.X.dot(self.X.T)` / `Y[:, :D].dot(self.X.T)`
// (vertex_histogram.py:177-179), the per-level `np.sum` (weisfeiler_lehman.py:270)
// and `np.dot(phi_x, phi_x.T)` (shortest_path.py:404) of the reference.
//
// Exactness: panel entries are integer counts <= 256 (exact in bf16), products and
// partial sums are integers < 2^24 (exact in fp32) -- the host checks both bounds
// before choosing this path -- so the result is bit-identical to the reference's
// float64 arithmetic.
//
// Kernel shape (one persistent CTA per SM, 256 threads):
//   warp 0      TMA producer   (one elected lane)
//   warp 1      MMA issuer     (one elected lane; UMMA 128 x 256 x 16, bf16 -> f32)
//   warp 2      TMEM allocator (512 columns = two 128x256 fp32 accumulators)
//   warps 4..7  epilogue       (tcgen05.ld -> diagonal fix / normalise -> global, plus
//                               the mirrored tile for the symmetric case)
// smem: 4 stages x (A 128x64 + B 256x64 bf16) = 192 KiB, mbarrier ring.
#pragma once
#include "common.cuh"

namespace gk {

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BK * 2;
constexpr int B_BYTES = BN * BK * 2;
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int GEMM_THREADS = 256;
// epilogue staging for TMA stores: per epilogue warp two 32x32 fp32 tiles (double buffer)
constexpr int EPI_BUF_BYTES = 32 * 32 * 4;
constexpr int EPI_BYTES = 4 * 2 * EPI_BUF_BYTES;
constexpr int GEMM_SMEM = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + 1024 /*align*/ + EPI_BYTES;

struct GramParams {
  const int2* tiles;   // {first A row, first B row} in panel-row (graph) coordinates
  int n_tiles;
  int num_k_blocks;    // Dc_pad / 64
  int a_row_end;       // rows >= a_row_end are not stored
  int b_row_end;
  int c_row0, c_col0;  // K[row - c_row0][col - c_col0]
  void* out;
  long long ld;
  int mirror;          // also store the transposed element (symmetric case): 1 = row stores, 2 = TMA store of a transposed tile
  int fix_diag;        // K[g][g] = diag[g] (self similarity over ALL columns)
  int nan_to_num;
  int vec_ok;          // 32-byte aligned rows: 256-bit vector stores allowed
  const double* diag;  // self similarity per graph (fp64, exact integers)
  long long* prof;     // optional [gridDim.x][8] cycle counters (GRAKEL_B200_PROF), else NULL
  int tma_store;       // fp32 output through a TMA store of smem-staged 32x32 blocks (tmC valid)
  // multi-GPU row tiling (gk_comm_init): the mirrored half of a tile belongs to the rank that owns K rows
  // [tile.y, tile.y + 256) and is stored straight into THAT rank's row block over NVLink (peer-mapped memory);
  // peer[r] = first row of rank r's block, peer_rows (a multiple of 256) rows per rank.  n_peers = 0: one GPU.
  void* peer[8];
  int n_peers;
  int peer_rows;
  // split-precision GEMM of real-valued features (ShortestPathAttr): k-blocks [k_block0, k_block0 + num_k_blocks) of
  // the panels, result ADDED to `out` when accumulate != 0 (the k range is processed in chunks so that no fp32
  // accumulator sums more than a few hundred MMAs)
  int accumulate;
  int k_block0;
  int k_chunk;         // k-blocks per accumulator (0 = all): the kernel folds every chunk of a tile into `out` itself
  int k_split;         // chunking applies to k-blocks [0, k_split); the rest is ONE chunk (0 = chunk everything)
  const int* nkb_dev;  // gram_tc2_kernel: number of k-blocks chosen ON THE DEVICE (asynchronous pass, features.cuh
                       // GramDyn), NULL = num_k_blocks
  // Tail fused into the epilogue (TMA-store path with the TMA mirror): the pair updates of the sparse tail, bucketed by
  // 32 x 32 block of K (tail_bucket, features.cuh: tb_cnt[block] entries {row | col << 5 | integer value << 10} at
  // tb_ent[block * TB_CAP ...]), are added to the staged block in shared memory before its bulk store -- instead of a
  // DRAM read-modify-write per update after the GEMM.  NULL = off.
  const unsigned* tb_cnt;
  const unsigned* tb_ent;
  int tb_nb32;         // blocks per row of K
  int tb_cap;          // entries per block the epilogue applies (<= TB_CAP, the storage stride; smaller only in tests)
  int tb_sym;          // 1: only pairs with row < column are bucketed (blocks on or above the block diagonal); a block below
                       // it reads its transpose's list with rows and columns swapped, a diagonal block applies both orders
};
constexpr int TB_CAP = 32;  // entries per block applied in the epilogue (one per lane); the rest go through tb overflow

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"((unsigned long long)map), "r"(bar), "r"(x), "r"(y)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int x, int y) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"((unsigned long long)map), "r"(src), "r"(x), "r"(y)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128-byte-swizzled shared-memory matrix descriptor (one 64-element bf16
// row = 128 B; 8-row groups 1024 B apart).  Bit layout: cute::UMMA::SmemDescriptor.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);  // start address   [0,14)
  d |= (uint64_t)1 << 16;                    // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset  [32,46)
  d |= (uint64_t)1 << 46;                    // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
  return d;
}

// instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=256
constexpr uint32_t UMMA_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                ((uint32_t)(BM >> 4) << 24);
// the same shape with A = B = tf32 (kind::tf32: 32 elements per 128-byte k-block, K = 8 per instruction)
constexpr uint32_t UMMA_IDESC_TF32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                     ((uint32_t)(BM >> 4) << 24);
constexpr int BK_TF32 = 32;

template <typename OutT, bool NORMALIZE>
__device__ __forceinline__ OutT epilogue_value(float acc, int arow, int bcol, double drow, double dcol,
                                               const GramParams& p) {
  if constexpr (NORMALIZE) {
    double v = (double)acc;
    if (p.fix_diag && arow == bcol) v = drow;
    v = v / sqrt(drow * dcol);  // np.divide(K, np.sqrt(np.outer(d, d)))  (fp64, correctly rounded)
    if (p.nan_to_num) {
      if (v != v) v = 0.0;
      else if (isinf(v)) v = v > 0 ? 1.7976931348623157e308 : -1.7976931348623157e308;
    }
    return (OutT)v;
  } else {
    if (p.fix_diag && arow == bcol) return (OutT)drow;
    return (OutT)acc;
  }
}

// Epilogue of one 128 x 256 accumulator for fp32, un-normalised output through TMA stores (shared by the
// one-CTA and the CTA-pair kernel).  `tile` = {first A row of THIS CTA's 128 rows, first B row}; `acc` = TMEM
// address of the accumulator's first column in this warp's lane quarter; `my_buf` = the warp's two staging tiles.
// tensor maps of the peers' row blocks (multi-GPU, mirrored blocks through TMA: GRAKEL_B200_DIST_TMA)
struct PeerMaps {
  CUtensorMap m[8];
};

__device__ __forceinline__ void epi_tma_store_tile(const GramParams& p, const CUtensorMap* tmC_ptr, int2 tile,
                                                   uint32_t acc, int ew, int lane, uint32_t my_buf,
                                                   const CUtensorMap* peer_maps = nullptr) {
  using OutT = float;
  const CUtensorMap& tmC = *tmC_ptr;
  OutT* __restrict__ out = reinterpret_cast<OutT*>(p.out);
  const int row = ew * 32 + lane;
  const int arow = tile.x + row;
  const bool row_ok = arow < p.a_row_end;
  const bool diag_tile = p.fix_diag && (tile.x < tile.y + BN) && (tile.y < tile.x + BM);
  // Direct block: registers -> 128B-swizzled smem tile -> ONE TMA store per 32x32 block (the
  // SM->L2 path is bound by write requests: 1 bulk request instead of 128).  TMA clips at the
  // matrix edge, so partial tiles need no predicates.  Mirrored block: coalesced 128-byte
  // stores straight from registers (lanes = consecutive rows of the tile).
  const int arow0 = tile.x + ew * 32;
  float dself = 0.f;
  if (diag_tile && row_ok) dself = (float)p.diag[arow];
  OutT* mptr = out + (long long)tile.y * p.ld + arow;
  if (p.n_peers) {  // the mirrored block lives in its owner's row block (this rank's own for tiles of its diagonal block)
    const int owner = tile.y / p.peer_rows;
    mptr = reinterpret_cast<OutT*>(p.peer[owner]) + (long long)(tile.y - owner * p.peer_rows) * p.ld + arow;
  }
  const long long ld = p.ld;
  const bool rows_in = arow0 < p.a_row_end;  // warp-uniform: any row of this warp's block inside
#pragma unroll 1
  for (int c0 = 0; c0 < BN; c0 += 32) {
    uint32_t v[32];
    // fused tail: this block's and the mirrored block's update lists are fetched before the accumulator is (no
    // dependent chain: every lane loads "its" entry slot speculatively, the count says which lanes hold one)
    unsigned tn_d = 0, tn_m = 0;
    unsigned te_d = 0u, te_m = 0u;
    int swap_d = 0, swap_m = 0;  // 0: entry as stored, 1: rows and columns swapped, 2: both (diagonal block, symmetric lists)
    if (p.tb_cnt && rows_in && tile.y + c0 < p.b_row_end) {
      const int bx = arow0 >> 5, by = (tile.y + c0) >> 5;
      if (p.tb_sym) {  // one list serves the block and its mirror image
        const long long bl = (long long)min(bx, by) * p.tb_nb32 + max(bx, by);
        tn_d = tn_m = min(__ldg(&p.tb_cnt[bl]), (unsigned)p.tb_cap);
        te_d = te_m = __ldg(&p.tb_ent[bl * TB_CAP + lane]);
        swap_d = bx < by ? 0 : (bx > by ? 1 : 2);
        swap_m = bx < by ? 1 : (bx > by ? 0 : 2);
      } else {
        const long long bd = (long long)bx * p.tb_nb32 + by;
        const long long bm = (long long)by * p.tb_nb32 + bx;
        tn_d = min(__ldg(&p.tb_cnt[bd]), (unsigned)p.tb_cap);
        tn_m = min(__ldg(&p.tb_cnt[bm]), (unsigned)p.tb_cap);
        te_d = __ldg(&p.tb_ent[bd * TB_CAP + lane]);
        te_m = __ldg(&p.tb_ent[bm * TB_CAP + lane]);
      }
    }
    tc_ld32(acc + (uint32_t)c0, v);
    const int bcol0 = tile.y + c0;
    if (bcol0 >= p.b_row_end || !rows_in) { mptr += 32 * ld; continue; }  // warp-uniform
    if (diag_tile) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (arow == bcol0 + j) v[j] = __float_as_uint(dself);
    }
    // buffers: without the TMA mirror the two staging tiles alternate between column blocks; with it
    // tile 0 stages the direct block and tile 1 the transposed one.  Either way the bulk group that
    // last read the tile about to be overwritten is the second newest, hence wait_group.read 1.
    const uint32_t buf = p.mirror == 2 ? my_buf : my_buf + (uint32_t)((c0 >> 5) & 1) * EPI_BUF_BYTES;
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");  // this buffer's previous store
    __syncwarp();
    const uint32_t rowaddr = buf + (uint32_t)lane * 128u;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const uint32_t a = rowaddr + (uint32_t)((q ^ (lane & 7)) << 4);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v[4 * q]), "r"(v[4 * q + 1]),
                   "r"(v[4 * q + 2]), "r"(v[4 * q + 3])
                   : "memory");
    }
    if (tn_d) {  // warp-uniform
      __syncwarp();
      if ((unsigned)lane < tn_d) {
        const uint32_t e0 = te_d & 31u, e1 = (te_d >> 5) & 31u;
        const float val = (float)(te_d >> 10);
        if (swap_d != 1) {
          const uint32_t r = e0, c = e1;
          const uint32_t a = buf + r * 128u + ((((c >> 2) ^ (r & 7u))) << 4) + ((c & 3u) << 2);
          asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(a), "f"(val) : "memory");
        }
        if (swap_d != 0) {
          const uint32_t r = e1, c = e0;
          const uint32_t a = buf + r * 128u + ((((c >> 2) ^ (r & 7u))) << 4) + ((c & 3u) << 2);
          asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(a), "f"(val) : "memory");
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      tma_store_2d(&tmC, buf, bcol0 - p.c_col0, arow0 - p.c_row0);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    if (p.mirror == 2) {
      // Mirrored block K[bcol0 + r][arow0 + c] = v_c[r]: transposed into the second staging tile (lane c
      // writes column c; with the 128-byte swizzle the 32 lanes of one st.shared hit 32 distinct banks)
      // and stored by ONE bulk request instead of 32 row stores of 128 bytes.  TMA clips at the edges.
      const uint32_t mbuf = my_buf + EPI_BUF_BYTES;
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      __syncwarp();
      const uint32_t cbase = mbuf + (uint32_t)((lane & 3) << 2);
      const uint32_t cq = (uint32_t)(lane >> 2);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const uint32_t a = cbase + (uint32_t)j * 128u + ((cq ^ (uint32_t)(j & 7)) << 4);
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(v[j]) : "memory");
      }
      if (tn_m) {  // the mirrored block K[bcol0 + r][arow0 + c] has its own list (the updates are symmetric, both halves are bucketed)
        __syncwarp();
        if ((unsigned)lane < tn_m) {
          const uint32_t e0 = te_m & 31u, e1 = (te_m >> 5) & 31u;
          const float val = (float)(te_m >> 10);
          if (swap_m != 1) {
            const uint32_t r = e0, c = e1;
            const uint32_t a = mbuf + r * 128u + ((((c >> 2) ^ (r & 7u))) << 4) + ((c & 3u) << 2);
            asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(a), "f"(val) : "memory");
          }
          if (swap_m != 0) {
            const uint32_t r = e1, c = e0;
            const uint32_t a = mbuf + r * 128u + ((((c >> 2) ^ (r & 7u))) << 4) + ((c & 3u) << 2);
            asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(a), "f"(val) : "memory");
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        if (p.n_peers && peer_maps) {  // the block belongs to the rank that owns K rows [bcol0, bcol0 + 32): one bulk store over NVLink
          const int owner = tile.y / p.peer_rows;
          tma_store_2d(&peer_maps[owner], mbuf, arow0, bcol0 - owner * p.peer_rows);
        } else {
          tma_store_2d(&tmC, mbuf, arow0 - p.c_col0, bcol0 - p.c_row0);
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    } else if (p.mirror && row_ok) {
      if (bcol0 + 32 <= p.b_row_end) {
#pragma unroll
        for (int j = 0; j < 32; ++j) { *reinterpret_cast<uint32_t*>(mptr) = v[j]; mptr += ld; }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (bcol0 + j < p.b_row_end) *reinterpret_cast<uint32_t*>(mptr) = v[j];
          mptr += ld;
        }
      }
    } else {
      mptr += 32 * ld;
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // smem reusable by the next tile
  __syncwarp();
}

template <typename OutT, bool NORMALIZE, int KIND = 0 /* 0: bf16 operands, 1: tf32 operands */>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gram_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, GramParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  const uint32_t holder = bar_base + 8u * (2 * STAGES + 4);
  const uint32_t epi_base = (bar_base + 256u + 1023u) & ~1023u;  // 1024-byte aligned (128B swizzle atoms)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((unsigned long long)&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((unsigned long long)&tmB) : "memory");
    if (p.tma_store) asm volatile("prefetch.tensormap [%0];" ::"l"((unsigned long long)&tmC) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);  // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(holder), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(holder) : "memory");

  if (warp == 0) {
    if (lane == 0) {  // ---------------- TMA producer
      int stage = 0;
      uint32_t phase = 0;
      long long w_empty = 0;
      const long long t_start = clock64();
      for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
        const int2 tile = p.tiles[t];
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          const long long c0 = clock64();
          mbar_wait(empty_bar(stage), phase ^ 1u);
          w_empty += clock64() - c0;
          mbar_expect_tx(full_bar(stage), STAGE_BYTES);
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          const int kx = (p.k_block0 + kb) * (KIND == 1 ? BK_TF32 : BK);
          tma_load_2d(sa, &tmA, full_bar(stage), kx, tile.x);
          tma_load_2d(sa + A_BYTES, &tmB, full_bar(stage), kx, tile.y);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
      if (p.prof) { p.prof[blockIdx.x * 8 + 5] = w_empty; p.prof[blockIdx.x * 8 + 6] = clock64() - t_start; }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ---------------- MMA issuer
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      long long w_tempty = 0, w_full = 0;
      const long long t_start = clock64();
      const int kc_base = p.k_chunk > 0 ? p.k_chunk : p.num_k_blocks;
      const int ksplit = p.k_split > 0 ? p.k_split : p.num_k_blocks;
      for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x)
      for (int kb0 = 0, kc = 0; kb0 < p.num_k_blocks; kb0 += kc, ++it) {  // one accumulator per (tile, k-chunk)
        kc = kb0 < ksplit ? min(kc_base, ksplit - kb0) : p.num_k_blocks - kb0;
        const int as = it & 1;
        const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
        long long c0 = clock64();
        mbar_wait(tempty_bar(as), aphase ^ 1u);
        w_tempty += clock64() - c0;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
        const int kb1 = min(p.num_k_blocks, kb0 + kc);
        for (int kb = kb0; kb < kb1; ++kb) {
          c0 = clock64();
          mbar_wait(full_bar(stage), phase);
          w_full += clock64() - c0;
          tc_fence_after();
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          const uint64_t adesc = umma_desc_sw128(sa);
          const uint64_t bdesc = umma_desc_sw128(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // +32 bytes along K inside the 128-byte swizzle row = +2 in the address field
            if constexpr (KIND == 1)
              tc_mma_tf32(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), UMMA_IDESC_TF32,
                          (uint32_t)(((kb - kb0) | k) != 0));
            else
              tc_mma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), UMMA_IDESC,
                          (uint32_t)(((kb - kb0) | k) != 0));
          }
          tc_commit(empty_bar(stage));  // frees the smem slot when these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        tc_commit(tfull_bar(as));  // accumulator complete
      }
      if (p.prof) { p.prof[blockIdx.x * 8 + 2] = w_tempty; p.prof[blockIdx.x * 8 + 3] = w_full; p.prof[blockIdx.x * 8 + 4] = clock64() - t_start; }
    }
  } else if (warp >= 4) {  // ---------------- epilogue
    const int ew = warp - 4;  // TMEM lanes [32*ew, 32*ew+32)
    const int row = ew * 32 + lane;
    OutT* __restrict__ out = reinterpret_cast<OutT*>(p.out);
    int it = 0;
    long long w_tfull = 0, t_work = 0;
    const int kc_base = p.k_chunk > 0 ? p.k_chunk : p.num_k_blocks;
    const int ksplit = p.k_split > 0 ? p.k_split : p.num_k_blocks;
    for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x)
    for (int kb0 = 0, kc_e = 0; kb0 < p.num_k_blocks; kb0 += kc_e, ++it) {
      kc_e = kb0 < ksplit ? min(kc_base, ksplit - kb0) : p.num_k_blocks - kb0;
      const int as = it & 1;
      const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
      const int2 tile = p.tiles[t];
      const bool acc_out = p.accumulate || kb0 > 0;  // later chunks of the tile add to what the earlier ones stored
      const long long c_w = clock64();
      mbar_wait(tfull_bar(as), aphase);
      const long long c_s = clock64();
      w_tfull += c_s - c_w;
      tc_fence_after();
      const int arow = tile.x + row;
      const bool row_ok = arow < p.a_row_end;
      // Tile-level classification keeps per-element branches out of the common case: the
      // epilogue is issue-bound (4 warps), so the interior path is pure tcgen05.ld + stores.
      const bool interior = (tile.x + BM <= p.a_row_end) && (tile.y + BN <= p.b_row_end);
      const bool diag_tile = p.fix_diag && (tile.x < tile.y + BN) && (tile.y < tile.x + BM);
      if (sizeof(OutT) == 4 && !NORMALIZE && p.tma_store) {
        epi_tma_store_tile(p, &tmC, tile, tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(as * BN), ew, lane,
                           epi_base + (uint32_t)ew * (2u * EPI_BUF_BYTES));
      } else if (KIND == 0 && !NORMALIZE && interior && !diag_tile && p.vec_ok) {
        OutT* drow_ptr = out + (long long)(arow - p.c_row0) * p.ld + (tile.y - p.c_col0);
        OutT* mptr = out + (long long)tile.y * p.ld + arow;  // mirror: K[col][row], one row of K per tile column
        const long long ld = p.ld;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t v[32];
          tc_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(as * BN + c0), v);
          if constexpr (sizeof(OutT) == 4) {
            // 256-bit stores (one full 32-byte sector per lane): the SM->L2 path is bound by the
            // number of write requests, not bytes, so halving the request count matters
#pragma unroll
            for (int j = 0; j < 32; j += 8)
              asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(drow_ptr + c0 + j),
                           "r"(v[j]), "r"(v[j + 1]), "r"(v[j + 2]), "r"(v[j + 3]), "r"(v[j + 4]), "r"(v[j + 5]),
                           "r"(v[j + 6]), "r"(v[j + 7])
                           : "memory");
            if (p.mirror) {
#pragma unroll
              for (int j = 0; j < 32; ++j) { *reinterpret_cast<uint32_t*>(mptr) = v[j]; mptr += ld; }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const double d0 = (double)__uint_as_float(v[j]), d1 = (double)__uint_as_float(v[j + 1]),
                           d2 = (double)__uint_as_float(v[j + 2]), d3 = (double)__uint_as_float(v[j + 3]);
              asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(drow_ptr + c0 + j), "d"(d0), "d"(d1),
                           "d"(d2), "d"(d3)
                           : "memory");
            }
            if (p.mirror) {
#pragma unroll
              for (int j = 0; j < 32; ++j) { *mptr = (OutT)__uint_as_float(v[j]); mptr += ld; }
            }
          }
        }
      } else {
        double drow = 0.0;
        if ((NORMALIZE || p.fix_diag) && row_ok) drow = p.diag[arow];
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t v[32];
          tc_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(as * BN + c0), v);
          const int bcol0 = tile.y + c0;
          if (bcol0 >= p.b_row_end) continue;  // warp-uniform
          double dcol_l = 0.0;
          if (NORMALIZE) {
            const int bc = bcol0 + lane;
            dcol_l = bc < p.b_row_end ? p.diag[bc] : 1.0;
          }
          OutT vals[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            double dcol = 0.0;
            if (NORMALIZE) dcol = __shfl_sync(0xffffffffu, dcol_l, j);
            vals[j] = epilogue_value<OutT, NORMALIZE>(__uint_as_float(v[j]), arow, bcol0 + j, drow, dcol, p);
          }
          if (row_ok) {
            OutT* dst = out + (long long)(arow - p.c_row0) * p.ld + (bcol0 - p.c_col0);
            if constexpr (sizeof(OutT) == 8) {
              if (!p.mirror && bcol0 + 32 <= p.b_row_end && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
                // full block of a split-precision GEMM: 16-byte loads / stores of the running fp64 sums
                double2* d2 = reinterpret_cast<double2*>(dst);
                if (acc_out) {
#pragma unroll
                  for (int j = 0; j < 16; ++j) { const double2 o = d2[j]; vals[2 * j] += o.x; vals[2 * j + 1] += o.y; }
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) d2[j] = make_double2(vals[2 * j], vals[2 * j + 1]);
                continue;
              }
            }
            if (acc_out) {  // later k-chunks of a split-precision GEMM: this CTA owns the tile, no atomics needed
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (bcol0 + j < p.b_row_end) vals[j] += dst[j];
            }
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (bcol0 + j < p.b_row_end) dst[j] = vals[j];
            if (p.mirror) {  // lanes hold consecutive rows -> coalesced transposed store
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (bcol0 + j < p.b_row_end) out[(long long)(bcol0 + j) * p.ld + arow] = vals[j];
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));
      t_work += clock64() - c_s;
    }
    if (p.prof && warp == 4 && lane == 0) { p.prof[blockIdx.x * 8 + 0] = w_tfull; p.prof[blockIdx.x * 8 + 1] = t_work; }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all TMA stores landed
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---------------------------------------------------------------------------
// Exact CUDA-core Gram over a u32 count panel (int64 accumulation).  Used (a) as
// the on-device cross-check of the tensor-core kernel in the tests and (b) when a
// count exceeds 256 or a Gram value exceeds 2^24, where bf16 x bf16 -> fp32 would
// no longer be exact.
template <typename OutT, bool NORMALIZE>
__global__ void __launch_bounds__(256)
gram_simt_kernel(const unsigned* __restrict__ panel, long long ldp, int kdim, int a_row0, int a_row_end,
                 int b_row0, int b_row_end, GramParams p) {
  __shared__ unsigned As[16][17], Bs[16][17];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int arow = a_row0 + blockIdx.y * 16 + ty;
  const int bcol = b_row0 + blockIdx.x * 16 + tx;
  unsigned long long acc = 0;
  for (int k0 = 0; k0 < kdim; k0 += 16) {
    const int ar = a_row0 + blockIdx.y * 16 + ty;
    const int br = b_row0 + blockIdx.x * 16 + ty;
    As[ty][tx] = (ar < a_row_end && k0 + tx < kdim) ? panel[(long long)ar * ldp + k0 + tx] : 0u;
    Bs[ty][tx] = (br < b_row_end && k0 + tx < kdim) ? panel[(long long)br * ldp + k0 + tx] : 0u;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) acc += (unsigned long long)As[ty][k] * Bs[tx][k];
    __syncthreads();
  }
  if (arow >= a_row_end || bcol >= b_row_end) return;
  double v = (double)acc;
  const double drow = p.diag[arow], dcol = p.diag[bcol];
  if (p.fix_diag && arow == bcol) v = drow;
  if (NORMALIZE) {
    v = v / sqrt(drow * dcol);
    if (p.nan_to_num) {
      if (v != v) v = 0.0;
      else if (isinf(v)) v = v > 0 ? 1.7976931348623157e308 : -1.7976931348623157e308;
    }
  }
  reinterpret_cast<OutT*>(p.out)[(long long)(arow - p.c_row0) * p.ld + (bcol - p.c_col0)] = (OutT)v;
}

// K with no contracted columns: zero matrix (+ exact diagonal / normalisation)
template <typename OutT, bool NORMALIZE>
__global__ void __launch_bounds__(256)
gram_empty_kernel(int a_row0, int a_row_end, int b_row0, int b_row_end, GramParams p) {
  const long long n_cols = b_row_end - b_row0;
  const long long total = (long long)(a_row_end - a_row0) * n_cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int arow = a_row0 + (int)(i / n_cols), bcol = b_row0 + (int)(i % n_cols);
    double v = 0.0;
    const double drow = p.diag[arow], dcol = p.diag[bcol];
    if (p.fix_diag && arow == bcol) v = drow;
    if (NORMALIZE) {
      v = v / sqrt(drow * dcol);
      if (p.nan_to_num) {
        if (v != v) v = 0.0;
        else if (isinf(v)) v = v > 0 ? 1.7976931348623157e308 : -1.7976931348623157e308;
      }
    }
    reinterpret_cast<OutT*>(p.out)[(long long)(arow - p.c_row0) * p.ld + (bcol - p.c_col0)] = (OutT)v;
  }
}

}  // namespace gk
