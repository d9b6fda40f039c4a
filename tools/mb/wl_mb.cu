// Micro-benchmark: where does the time of the WL signature / insert phase go?
// Variants of the thread-per-vertex signature kernel with pieces disabled, plus raw
// atomic-throughput probes on a hash table of the real size.  Build: tools/mb/build.sh
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <random>
#include "../../grakel_b200/csrc/wl.cuh"
using namespace gk;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int MODE>  // 0 full, 1 no insert, 2 no insert + no sig store, 3 only row_ptr/col_idx loads
__global__ void __launch_bounds__(256) sig_variant(int V, const int* __restrict__ row_ptr, const int* __restrict__ col_idx,
    const int* __restrict__ lab_in, int* __restrict__ sig_nbr, unsigned long long seed, unsigned long long* ht_keys,
    int* ht_rep, unsigned ht_mask, int* __restrict__ slot_of) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  const int beg = row_ptr[v];
  const int deg = row_ptr[v + 1] - beg;
  if (deg > 8) return;
  int nb[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) nb[i] = i < deg ? col_idx[beg + i] : -1;
  int xs[8];
  if (MODE == 3) {
#pragma unroll
    for (int i = 0; i < 8; ++i) xs[i] = nb[i];
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) xs[i] = nb[i] >= 0 ? lab_in[nb[i]] : 0x7fffffff;
  }
  const int own = lab_in[v];
  GK_CSWAP(xs[0], xs[1]) GK_CSWAP(xs[2], xs[3]) GK_CSWAP(xs[4], xs[5]) GK_CSWAP(xs[6], xs[7])
  GK_CSWAP(xs[0], xs[2]) GK_CSWAP(xs[1], xs[3]) GK_CSWAP(xs[4], xs[6]) GK_CSWAP(xs[5], xs[7])
  GK_CSWAP(xs[1], xs[2]) GK_CSWAP(xs[5], xs[6]) GK_CSWAP(xs[0], xs[4]) GK_CSWAP(xs[3], xs[7])
  GK_CSWAP(xs[1], xs[5]) GK_CSWAP(xs[2], xs[6])
  GK_CSWAP(xs[1], xs[4]) GK_CSWAP(xs[3], xs[6])
  GK_CSWAP(xs[2], xs[4]) GK_CSWAP(xs[3], xs[5])
  GK_CSWAP(xs[3], xs[4])
  unsigned long long t = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (i < deg) {
      if (MODE < 2) sig_nbr[beg + i] = xs[i];
      t += sig_term(xs[i], i, seed);
    }
  }
  const unsigned long long key = sig_final(t, own, deg, seed);
  if (MODE == 0) slot_of[v] = (int)ht_insert(ht_keys, ht_rep, ht_mask, key, v);
  else slot_of[v] = (int)key;
}

// raw probes: every thread touches one pseudo-random slot
template <int OP>
__global__ void __launch_bounds__(256) probe(int n, unsigned long long* k64, unsigned* k32, unsigned mask, int* out, int distinct) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long key = mix64((unsigned long long)i + 12345ULL) >> 1;
  const unsigned slot = (unsigned)(key * 0x9E3779B97F4A7C15ULL >> 20) & mask;
  int r = 0;
  if (OP == 0) r = (int)__ldcg(&k64[slot]);                                            // scattered 8-byte load
  if (OP == 1) r = (int)atomicCAS(&k64[slot], EMPTY64, key);                           // CAS 64
  if (OP == 2) r = (int)atomicCAS(&k32[slot], 0xFFFFFFFFu, (unsigned)key);             // CAS 32
  if (OP == 3) atomicMin(reinterpret_cast<int*>(&k32[slot]), i);                        // RED min 32
  if (OP == 4) r = (int)atomicAdd(&k32[slot], 1u);                                     // ATOM add 32 with return
  if (OP == 5) k64[slot] = key;                                                        // scattered 8-byte store
  if (OP == 6) { unsigned long long p = __ldcg(&k64[slot]); if (p == EMPTY64) p = atomicCAS(&k64[slot], EMPTY64, key); r = (int)p; }
  if (OP == 7) atomicAdd(&k32[slot], 1u);                                              // RED add 32
  if (r == 0x12345) out[0] = r;
}

int main() {
  const int N = 10000, V0 = 40;
  std::mt19937 rng(1);
  std::vector<int> rp(1, 0), ci, lab;
  int V = 0;
  for (int g = 0; g < N; ++g) {
    int n = 20 + rng() % 41;
    std::vector<std::vector<int>> adj(n);
    double p = 4.0 / (n - 1);
    for (int a = 0; a < n; ++a) for (int b = a + 1; b < n; ++b) if ((rng() % 100000) < p * 100000) { adj[a].push_back(b); adj[b].push_back(a); }
    for (int a = 0; a < n; ++a) { for (int w : adj[a]) ci.push_back(V + w); rp.push_back((int)ci.size()); lab.push_back(rng() % 380000); }
    V += n;
  }
  const int E = (int)ci.size();
  printf("V %d E %d\n", V, E);
  int *d_rp, *d_ci, *d_lab, *d_sig, *d_slot, *d_rep, *d_out;
  unsigned long long* d_keys; unsigned* d_k32;
  const size_t cap = 1 << 20;
  CK(cudaMalloc(&d_rp, (V + 1) * 4)); CK(cudaMalloc(&d_ci, E * 4)); CK(cudaMalloc(&d_lab, V * 4)); CK(cudaMalloc(&d_sig, E * 4));
  CK(cudaMalloc(&d_slot, V * 4)); CK(cudaMalloc(&d_rep, cap * 4)); CK(cudaMalloc(&d_keys, cap * 8)); CK(cudaMalloc(&d_k32, cap * 4));
  CK(cudaMalloc(&d_out, 64));
  CK(cudaMemcpy(d_rp, rp.data(), (V + 1) * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_ci, ci.data(), E * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_lab, lab.data(), V * 4, cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int nb = (V + 255) / 256;
  auto reset = [&]() { CK(cudaMemset(d_keys, 0xFF, cap * 8)); CK(cudaMemset(d_rep, 0x7F, cap * 4)); CK(cudaMemset(d_k32, 0xFF, cap * 4)); };
  auto timeit = [&](const char* name, auto launch) {
    float best = 1e9f;
    for (int it = 0; it < 5; ++it) {
      reset();
      CK(cudaDeviceSynchronize());
      cudaEventRecord(e0); launch(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
    }
    CK(cudaGetLastError());
    printf("%-44s %8.2f us\n", name, best * 1e3f);
  };
  timeit("empty kernel (1 launch)", [&]() { probe<9><<<nb, 256>>>(V, d_keys, d_k32, (unsigned)(cap - 1), d_out, 1); });
  timeit("sig full (thread8)", [&]() { sig_variant<0><<<nb, 256>>>(V, d_rp, d_ci, d_lab, d_sig, 7, d_keys, d_rep, (unsigned)(cap - 1), d_slot); });
  timeit("sig without insert", [&]() { sig_variant<1><<<nb, 256>>>(V, d_rp, d_ci, d_lab, d_sig, 7, d_keys, d_rep, (unsigned)(cap - 1), d_slot); });
  timeit("sig without insert, without sig_nbr store", [&]() { sig_variant<2><<<nb, 256>>>(V, d_rp, d_ci, d_lab, d_sig, 7, d_keys, d_rep, (unsigned)(cap - 1), d_slot); });
  timeit("sig: no label gather either", [&]() { sig_variant<3><<<nb, 256>>>(V, d_rp, d_ci, d_lab, d_sig, 7, d_keys, d_rep, (unsigned)(cap - 1), d_slot); });
  const char* names[8] = {"ldcg 8B scattered", "CAS64", "CAS32", "RED min32", "ATOM add32 (ret)", "store 8B scattered", "ldcg + CAS64", "RED add32"};
  for (int rep = 0; rep < 2; ++rep) {
    const int n = rep == 0 ? V : 4 * V;
    printf("-- %d threads, distinct keys\n", n);
    const int g = (n + 255) / 256;
    timeit(names[0], [&]() { probe<0><<<g, 256>>>(n, d_keys, d_k32, (unsigned)(cap - 1), d_out, 1); });
    timeit(names[1], [&]() { probe<1><<<g, 256>>>(n, d_keys, d_k32, (unsigned)(cap - 1), d_out, 1); });
    timeit(names[2], [&]() { probe<2><<<g, 256>>>(n, d_keys, d_k32, (unsigned)(cap - 1), d_out, 1); });
    timeit(names[3], [&]() { probe<3><<<g, 256>>>(n, d_keys, d_k32, (unsigned)(cap - 1), d_out, 1); });
    timeit(names[4], [&]() { probe<4><<<g, 256>>>(n, d_keys, d_k32, (unsigned)(cap - 1), d_out, 1); });
    timeit(names[5], [&]() { probe<5><<<g, 256>>>(n, d_keys, d_k32, (unsigned)(cap - 1), d_out, 1); });
    timeit(names[6], [&]() { probe<6><<<g, 256>>>(n, d_keys, d_k32, (unsigned)(cap - 1), d_out, 1); });
    timeit(names[7], [&]() { probe<7><<<g, 256>>>(n, d_keys, d_k32, (unsigned)(cap - 1), d_out, 1); });
  }
  // HBM write bandwidth probes (K store floor): memset of 400 MB
  {
    void* big; CK(cudaMalloc(&big, 400u << 20));
    for (int it = 0; it < 3; ++it) {
      cudaEventRecord(e0); CK(cudaMemsetAsync(big, 0, 400u << 20)); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      printf("memset 400 MiB: %.1f us = %.0f GB/s\n", ms * 1e3f, (400u << 20) / (ms * 1e-3) / 1e9);
    }
  }
  return 0;
}
