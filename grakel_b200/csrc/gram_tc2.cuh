// CTA-pair variant of the Gram GEMM: tcgen05.mma.cta_group::2, one 256 x 256 output tile per cluster of two
// CTAs (two SMs of one TPC), fp32 un-normalised output through the TMA-store epilogue of gram_tc.cuh.
//
// Why: the one-CTA kernel stalls on operand delivery (profiles/r01f_gram_tc_role_counters.txt: its MMA issuer
// waits on `full` 35 % of the time in dense mode, 45 % in head mode).  A 128 x 256 tile pulls 48 KB from L2 per
// 4.2 Mflop; at the tensor rate that is ~17 TB/s over the chip.  With cta_group::2 each CTA stages its own 128
// rows of A and only HALF of the B tile (the MMA reads the other half from the peer's shared memory): 32 KB per
// CTA for the same flops -- a third less L2 -> SM traffic -- and six 32 KB stages fit where four 48 KB ones did.
//
// Protocol (same shared-memory offsets in both CTAs; rank 0 = leader):
//   full[s]   (leader's only)  2 arrivals: the leader's arrive.expect_tx(2 x 32 KB) and the peer's remote arrive;
//             both CTAs' TMA loads complete their bytes on the LEADER's barrier (cp.async.bulk.tensor .cta_group::2)
//   empty[s]  (one per CTA)    armed by the leader's tcgen05.commit.cta_group::2 ... multicast::cluster 0b11
//   tfull[a]  (one per CTA)    same multicast commit after the tile's last MMA: each CTA's epilogue warps drain
//             their own 128 TMEM lanes
//   tempty[a] (leader's only)  8 arrivals: four epilogue warps of each CTA (the peer's arrive through mapa)
// Only the leader's warp 1 issues MMAs; producers and epilogues run in both CTAs.
#pragma once
#include "gram_tc.cuh"

namespace gk {

constexpr int BM2 = 256;  // rows of A per cluster tile (128 per CTA)
constexpr int STAGES2 = 6;
constexpr int A2_BYTES = BM * BK * 2;         // this CTA's 128 rows of A
constexpr int B2_BYTES = (BN / 2) * BK * 2;   // this CTA's half (128 rows) of the B tile
constexpr int STAGE2_BYTES = A2_BYTES + B2_BYTES;
constexpr int GEMM2_SMEM = STAGES2 * STAGE2_BYTES + 1024 /*align*/ + 256 /*barriers*/ + 1024 /*align*/ + EPI_BYTES;
static_assert(GEMM2_SMEM <= 232448, "shared memory budget of one CTA");

// instruction descriptor: D=f32, A=B=bf16, both K-major, M=256 (two CTAs), N=256
constexpr uint32_t UMMA2_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                 ((uint32_t)(BM2 >> 4) << 24);

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in the CTA of rank `rank`
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load into THIS CTA's shared memory whose bytes complete on a barrier that may live in the peer CTA
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"((unsigned long long)map), "r"(bar_cluster), "r"(x), "r"(y)
      : "memory");
}
__device__ __forceinline__ void tc2_commit_mc(uint32_t bar) {  // arrive on `bar` in both CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((unsigned short)3)
               : "memory");
}
__device__ __forceinline__ void tc2_mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gram_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmC, GramParams p,
                const __grid_constant__ PeerMaps pm) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES2 * STAGE2_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES2 + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES2 + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES2 + 2 + a); };
  const uint32_t holder = bar_base + 8u * (2 * STAGES2 + 4);
  const uint32_t epi_base = (bar_base + 256u + 1023u) & ~1023u;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;
  const int nkb = p.nkb_dev ? __ldg(p.nkb_dev) : p.num_k_blocks;  // >= 1 by construction (gram_choose)

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((unsigned long long)&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((unsigned long long)&tmC) : "memory");
    for (int s = 0; s < STAGES2; ++s) {
      mbar_init(full_bar(s), 2);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 8);  // four epilogue warps in each of the two CTAs
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {  // the same warp in both CTAs, same destination offset
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(holder), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised before anything arrives on them
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(holder) : "memory");

  if (warp == 0) {
    if (lane == 0) {  // ---------------- TMA producer (both CTAs)
      int stage = 0;
      uint32_t phase = 0;
      for (int t = cid; t < p.n_tiles; t += ncl) {
        const int2 tile = p.tiles[t];
        const int arow = tile.x + (int)rank * BM, brow = tile.y + (int)rank * (BN / 2);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t lfull = mapa_u32(full_bar(stage), 0);
          if (rank == 0) mbar_expect_tx(full_bar(stage), 2 * STAGE2_BYTES);
          else mbar_arrive_cluster(lfull);
          const uint32_t sa = smem_base + stage * STAGE2_BYTES;
          tma2_load_2d(sa, &tmA, lfull, kb * BK, arow);
          tma2_load_2d(sa + A2_BYTES, &tmA, lfull, kb * BK, brow);
          if (++stage == STAGES2) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {  // ---------------- MMA issuer (leader only)
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = cid; t < p.n_tiles; t += ncl, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
        mbar_wait(tempty_bar(as), aphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * STAGE2_BYTES;
          const uint64_t adesc = umma_desc_sw128(sa);
          const uint64_t bdesc = umma_desc_sw128(sa + A2_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            tc2_mma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), UMMA2_IDESC,
                         (uint32_t)((kb | k) != 0));
          tc2_commit_mc(empty_bar(stage));  // frees the slot in BOTH CTAs when these MMAs retire
          if (++stage == STAGES2) { stage = 0; phase ^= 1u; }
        }
        tc2_commit_mc(tfull_bar(as));  // accumulator complete: both CTAs' epilogues may drain their lanes
      }
    }
  } else if (warp >= 4) {  // ---------------- epilogue (both CTAs, own 128 TMEM lanes)
    const int ew = warp - 4;
    const uint32_t my_buf = epi_base + (uint32_t)ew * (2u * EPI_BUF_BYTES);
    int it = 0;
    for (int t = cid; t < p.n_tiles; t += ncl, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
      int2 tile = p.tiles[t];
      tile.x += (int)rank * BM;
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      epi_tma_store_tile(p, &tmC, tile, tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(as * BN), ew, lane, my_buf,
                         p.mirror == 2 && p.n_peers ? pm.m : nullptr);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(tempty_bar(as), 0));
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all TMA stores landed
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // neither CTA frees tensor memory (or exits) while the pair still works
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace gk
