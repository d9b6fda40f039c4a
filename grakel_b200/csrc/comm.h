// Multi-GPU plumbing inside the C-ABI (included by api.cu only): one process per GPU, one gk_handle per
// process, NCCL for the control plane (unique-id bootstrap, handle exchange, barriers) and for the final
// all-gather of BASELINE config 4; the data path of the row-tiled Gram is the GEMM epilogue itself, which
// stores the mirrored half of every tile into the owning rank's row block through peer-mapped (CUDA IPC)
// memory over NVLink (gram_tc.cuh, GramParams::peer).
//
// NCCL is resolved at run time (dlopen): a copy already loaded into the process (PyTorch bundles one) is
// re-used, otherwise the system library; single-GPU users never touch it.
#pragma once
#include <dlfcn.h>
#include <nccl.h>  // types and enums only; every entry point is looked up with dlsym

namespace gk {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
};

inline NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* env = getenv("GRAKEL_B200_NCCL_LIB");
    void* lib = nullptr;
    if (env && *env) lib = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL | RTLD_NOLOAD);  // already in the process
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) { api.error = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : ""); return; }
    api.lib = lib;
    auto sym = [&](const char* n) -> void* {
      void* p = dlsym(lib, n);
      if (!p && api.error.empty()) api.error = std::string("NCCL symbol missing: ") + n;
      return p;
    };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
  });
  return &api;
}

#define GK_NCCL(api, expr)                                                                                   \
  do {                                                                                                       \
    ncclResult_t _r = (expr);                                                                                \
    if (_r != ncclSuccess)                                                                                   \
      return gk::fail(GK_ERR_CUDA, std::string(#expr) + ": " + ((api)->GetErrorString ? (api)->GetErrorString(_r) : "NCCL error")); \
  } while (0)

constexpr int DIST_MAX_RANKS = 8;
constexpr int DIST_ALIGN = 256;  // row blocks are multiples of the CTA-pair tile

struct Comm {
  ncclComm_t comm = nullptr;
  int nranks = 1, rank = 0;
  int* d_token = nullptr;        // barrier payload
  unsigned char* d_handles = nullptr;  // [nranks + 1][64] IPC handles (slot nranks = ours)
  void* peer_base[DIST_MAX_RANKS] = {};  // peers' K buffers as mapped into this process (ours: h->K.p)
  void* exported = nullptr;      // the local allocation the peers currently map
  size_t exported_cap = 0;
};

// rows per rank: ceil(n / nranks) rounded up to the tile size (trailing ranks may own fewer or no rows)
inline long long dist_rows_per_rank(long long n, int nranks) {
  const long long per = (n + nranks - 1) / nranks;
  return (per + DIST_ALIGN - 1) / DIST_ALIGN * DIST_ALIGN;
}

// Tiles (256 x 256, {first K row, first K column}) rank `rank` computes for its row block: the upper triangle
// of its diagonal block, and of every rectangle it shares with another rank the tiles the parity rule gives it
// (global tile indices ti, tj: the owner of the smaller index computes the tile when ti + tj is even, the other
// one when it is odd).  The mirrored half of each tile is stored into the owner of rows [col, col + 256).
inline void dist_tiles(long long n, int nranks, int rank, std::vector<int2>& tiles) {
  const long long per = dist_rows_per_rank(n, nranks);
  const long long rb = std::min(n, rank * per), re = std::min(n, (rank + 1) * per);
  if (rb >= re) return;  // a trailing rank without rows
  const long long n_t = (n + DIST_ALIGN - 1) / DIST_ALIGN;
  // Bands of row tiles, column-major inside a band (the tiles in flight share panel rows in L2).  The column tiles
  // are visited owner by owner in ROTATED order -- rank r starts with its own block, then the blocks of r+1, r+2, ...
  // -- so that at any moment the ranks store their mirrored halves into DIFFERENT peers (a permutation of the NVLink
  // ports).  With every rank walking the owners 0, 1, 2, ... all of them wrote into the same GPU at the same time and
  // shared its ingress: 1.4 ms for the GEMM at 8 GPUs against 0.2 ms on one (profiles/r02q_bench8.json).
  const long long ti0 = rb / DIST_ALIGN, ti1 = (re + DIST_ALIGN - 1) / DIST_ALIGN;
  const long long BAND = 6;
  const long long t_per = per / DIST_ALIGN;  // column tiles per owner
  for (long long m0 = ti0; m0 < ti1; m0 += BAND) {
    const long long m1 = std::min(ti1, m0 + BAND);
    for (int step = 0; step < nranks; ++step) {
      const int owner = (rank + step) % nranks;
      for (long long tj = owner * t_per; tj < std::min(n_t, (owner + 1) * t_per); ++tj) {
        for (long long ti = m0; ti < m1; ++ti) {
          bool mine;
          if (owner == rank) mine = tj >= ti;
          else mine = (((ti + tj) & 1) == 0) == (ti < tj);
          if (mine) tiles.push_back(make_int2((int)(ti * DIST_ALIGN), (int)(tj * DIST_ALIGN)));
        }
      }
    }
  }
}

}  // namespace gk
