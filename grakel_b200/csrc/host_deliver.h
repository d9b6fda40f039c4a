// Host side of result delivery (included by api.cu only).
//
// The reference returns a fresh float64 ndarray on the host (kernel.py:167-204); for BASELINE config 2 that is
// 800 MB, and a plain fp64 D2H copy (14-15 ms of PCIe) was 92 % of the end-to-end call.  K entries are exact
// integers below 2^24 on the tensor-core path, so K is produced and moved as fp32 and widened on the host:
//
//   deliver_tri   symmetric N x N result: only the upper triangle crosses PCIe (N^2/2 fp32 = a quarter of the
//                 fp64 bytes), in row bands of about equal area through a ring of pinned staging buffers.  For a
//                 band [r0, r1) the staging buffer holds columns [r0, N); workers own column ranges of it, widen
//                 their part into rows r0..r1 of the destination and write its transpose -- rows >= r1, columns
//                 r0..r1 -- with 8 x 8 register transposes, so every destination line is written exactly once
//                 with full-line streaming stores while the next bands are in flight.
//   deliver_rows  any row block (transform results, row tiles of a multi-GPU job): all columns cross as fp32;
//                 optional fp64 normalisation K_ij / sqrt(d_i d_j) (+ nan_to_num) during the widening -- the
//                 same IEEE operations as the reference's numpy expression (kernel.py:198-203).
//
// Workers are a persistent pool pinned to the NUMA node of the creating thread; one dispatch per delivery
// (workers follow the band sequence through two atomics), not one per band.
#pragma once
#include <atomic>
#include <sys/mman.h>

namespace {

class HostPool {
 public:
  explicit HostPool(int n) : stop_(false), gen_(0), pending_(0) {
    // Workers stay on the NUMA node of the thread that creates the pool (the one that also allocates and touches
    // the pinned staging buffers; remote-socket workers made the widening slower with every added thread on the
    // 2-socket GPU hosts), one worker per PHYSICAL core -- two workers on hyper-thread siblings turn into the
    // stragglers every band waits for -- and not on the caller's core, which spins in cudaEventSynchronize.
    std::vector<int> cpus = pick_cores(local_node_cpus());
    // several processes on one host (one per GPU): each takes its own slice of the node's cores
    int local_rank = 0;
    if (const char* e = getenv("LOCAL_RANK")) local_rank = std::max(0, atoi(e));
    const int first = cpus.empty() ? 0 : (int)(((long long)local_rank * n) % (long long)cpus.size());
    for (int i = 0; i < n; ++i) {
      th_.emplace_back([this, i] { run(i); });
      if (!cpus.empty()) {
        cpu_set_t set;
        CPU_ZERO(&set);
        // GRAKEL_B200_HOST_FLOAT=1: workers may run on ANY of the picked CPUs (one per physical core) instead of one
        // fixed CPU each -- a worker whose core is taken by somebody else (the caller's own spinning thread after a
        // migration, a proxy thread) is moved by the scheduler instead of waiting for a time slice
        static const bool floating = getenv("GRAKEL_B200_HOST_FLOAT") && atoi(getenv("GRAKEL_B200_HOST_FLOAT")) != 0;
        if ((int)cpus.size() >= n && !floating) CPU_SET(cpus[(first + i) % (int)cpus.size()], &set);
        else for (int c : cpus) CPU_SET(c, &set);
        pthread_setaffinity_np(th_.back().native_handle(), sizeof(set), &set);
      }
    }
  }
  static int read_int(const char* fmt, int cpu) {
    char path[160];
    snprintf(path, sizeof(path), fmt, cpu);
    FILE* f = fopen(path, "r");
    if (!f) return -1;
    int v = -1;
    if (fscanf(f, "%d", &v) != 1) v = -1;
    fclose(f);
    return v;
  }
  // one logical CPU per physical core of `list`, the caller's core excluded (empty list / no topology: unchanged)
  static std::vector<int> pick_cores(const std::vector<int>& list) {
    if (list.empty() || getenv("GRAKEL_B200_HOST_NO_PIN")) return list;
    const int me = sched_getcpu();
    const int my_core = me >= 0 ? read_int("/sys/devices/system/cpu/cpu%d/topology/core_id", me) : -1;
    const int my_pkg = me >= 0 ? read_int("/sys/devices/system/cpu/cpu%d/topology/physical_package_id", me) : -1;
    std::vector<std::pair<int, int>> seen;
    std::vector<int> out;
    for (int c : list) {
      const int core = read_int("/sys/devices/system/cpu/cpu%d/topology/core_id", c);
      const int pkg = read_int("/sys/devices/system/cpu/cpu%d/topology/physical_package_id", c);
      if (core < 0) return list;
      if (core == my_core && pkg == my_pkg) continue;
      if (std::find(seen.begin(), seen.end(), std::make_pair(pkg, core)) != seen.end()) continue;
      seen.emplace_back(pkg, core);
      out.push_back(c);
    }
    return out.empty() ? list : out;
  }
  // NUMA node the current CUDA device hangs off (its PCIe root), -1 if unknown
  static int gpu_numa_node() {
    int dev = 0;
    char bus[32] = {0};
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetPCIBusId(bus, sizeof(bus), dev) != cudaSuccess) return -1;
    for (char* c = bus; *c; ++c) *c = (char)tolower(*c);
    char path[128];
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE* f = fopen(path, "r");
    if (!f) return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    return node;
  }
  static std::vector<int> node_cpus(int node) {
    std::vector<int> list;
    char path[128];
    snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
    FILE* f = fopen(path, "r");
    if (!f) return list;
    char buf[4096] = {0};
    if (fgets(buf, sizeof(buf), f)) {
      for (char* tok = strtok(buf, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
        int a = 0, b = 0;
        if (sscanf(tok, "%d-%d", &a, &b) == 2) { for (int c = a; c <= b; ++c) list.push_back(c); }
        else if (sscanf(tok, "%d", &a) == 1) list.push_back(a);
      }
    }
    fclose(f);
    return list;
  }
  // CPUs of the node the GPU is attached to (its DMA lands in that node's memory without crossing the socket
  // interconnect); failing that, of the node the calling thread runs on
  static std::vector<int> local_node_cpus() {
    std::vector<int> out;
    if (const char* e = getenv("GRAKEL_B200_HOST_ANY_NODE"))
      if (atoi(e) != 0) return out;
    const int gnode = gpu_numa_node();
    if (gnode >= 0) {
      out = node_cpus(gnode);
      if (!out.empty()) return out;
    }
    const int cpu = sched_getcpu();
    if (cpu < 0) return out;
    for (int node = 0; node < 64; ++node) {
      char path[128];
      snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
      FILE* f = fopen(path, "r");
      if (!f) break;
      char buf[4096] = {0};
      if (!fgets(buf, sizeof(buf), f)) { fclose(f); continue; }
      fclose(f);
      std::vector<int> list;  // "0-15,64-79"
      for (char* tok = strtok(buf, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
        int a = 0, b = 0;
        if (sscanf(tok, "%d-%d", &a, &b) == 2) { for (int c = a; c <= b; ++c) list.push_back(c); }
        else if (sscanf(tok, "%d", &a) == 1) list.push_back(a);
      }
      if (std::find(list.begin(), list.end(), cpu) != list.end()) return list;
    }
    return out;
  }
  ~HostPool() {
    { std::lock_guard<std::mutex> l(m_); stop_ = true; ++gen_; }
    cv_.notify_all();
    for (auto& t : th_) t.join();
  }
  int size() const { return (int)th_.size(); }
  // start fn(worker_index) on every worker; wait() returns when all are done.  One delivery at a time.
  void start(const std::function<void(int)>& fn) {
    call_m_.lock();  // engines on different threads share the pool
    std::unique_lock<std::mutex> l(m_);
    fn_ = &fn;
    pending_ = (int)th_.size();
    ++gen_;
    cv_.notify_all();
  }
  void wait() {
    {
      std::unique_lock<std::mutex> l(m_);
      done_.wait(l, [this] { return pending_ == 0; });
    }
    call_m_.unlock();
  }
  void run_all(const std::function<void(int)>& fn) { start(fn); wait(); }

 private:
  void run(int idx) {
    unsigned long long seen = 0;
    for (;;) {
      const std::function<void(int)>* fn;
      {
        std::unique_lock<std::mutex> l(m_);
        cv_.wait(l, [&] { return gen_ != seen; });
        seen = gen_;
        if (stop_) return;
        fn = fn_;
      }
      (*fn)(idx);
      {
        std::lock_guard<std::mutex> l(m_);
        if (--pending_ == 0) done_.notify_all();
      }
    }
  }
  std::vector<std::thread> th_;
  std::mutex m_, call_m_;
  std::condition_variable cv_, done_;
  const std::function<void(int)>* fn_ = nullptr;
  bool stop_;
  unsigned long long gen_;
  int pending_;
};

HostPool& host_pool() {
  static HostPool pool([] {
    int n = std::min(16, (int)std::thread::hardware_concurrency() / 4);  // 16 measured best on the 2 x 32-core GPU hosts (profiles/r02c_*)
    // one process per GPU (torchrun sets LOCAL_WORLD_SIZE): the processes share the host's cores -- eight pools of 16
    // spinning workers on 64 cores took 687 ms for a delivery that takes 9 ms with the cores to itself (profiles/r02q_bench8.json)
    if (const char* e = getenv("LOCAL_WORLD_SIZE")) {
      const int lws = atoi(e);
      if (lws > 1) n = std::max(2, std::min(n, (int)std::thread::hardware_concurrency() / 2 / lws - 1));
    }
    if (const char* e = getenv("GRAKEL_B200_HOST_THREADS")) n = atoi(e);
    return std::max(1, std::min(n, 128));
  }());
  return pool;
}

// ---- row widening: n floats -> n doubles, optional normalisation by dr * dc[j]
__attribute__((target("avx2"))) inline void widen_row_avx2(const float* __restrict__ src, double* __restrict__ dst, long long n) {
  long long j = 0;
  for (; j < n && (reinterpret_cast<uintptr_t>(dst + j) & 31); ++j) dst[j] = (double)src[j];
  for (; j + 8 <= n; j += 8) {
    const __m256 f = _mm256_loadu_ps(src + j);
    _mm256_stream_pd(dst + j, _mm256_cvtps_pd(_mm256_castps256_ps128(f)));
    _mm256_stream_pd(dst + j + 4, _mm256_cvtps_pd(_mm256_extractf128_ps(f, 1)));
  }
  for (; j < n; ++j) dst[j] = (double)src[j];
}

inline double norm_one(double v, double dr, double dc, int nan_to_num) {
  v = v / std::sqrt(dr * dc);
  if (nan_to_num) {
    if (v != v) v = 0.0;
    else if (std::isinf(v)) v = v > 0 ? 1.7976931348623157e308 : -1.7976931348623157e308;
  }
  return v;
}

__attribute__((target("avx2"))) inline void widen_row_norm_avx2(const float* __restrict__ src, double* __restrict__ dst, long long n,
                                                                double dr, const double* __restrict__ dc, int nan_to_num) {
  long long j = 0;
  for (; j < n && (reinterpret_cast<uintptr_t>(dst + j) & 31); ++j) dst[j] = norm_one((double)src[j], dr, dc[j], nan_to_num);
  const __m256d vdr = _mm256_set1_pd(dr);
  const __m256d vmax = _mm256_set1_pd(1.7976931348623157e308);
  for (; j + 4 <= n; j += 4) {
    __m256d v = _mm256_cvtps_pd(_mm_loadu_ps(src + j));
    v = _mm256_div_pd(v, _mm256_sqrt_pd(_mm256_mul_pd(vdr, _mm256_loadu_pd(dc + j))));
    if (nan_to_num) {
      v = _mm256_and_pd(v, _mm256_cmp_pd(v, v, _CMP_ORD_Q));  // NaN -> +0.0
      v = _mm256_max_pd(_mm256_min_pd(v, vmax), _mm256_sub_pd(_mm256_setzero_pd(), vmax));  // +-inf -> +-DBL_MAX
    }
    _mm256_stream_pd(dst + j, v);
  }
  for (; j < n; ++j) dst[j] = norm_one((double)src[j], dr, dc[j], nan_to_num);
}

inline void widen_row(const float* __restrict__ src, double* __restrict__ dst, long long n) {
  static const bool avx2 = __builtin_cpu_supports("avx2");
  if (avx2) return widen_row_avx2(src, dst, n);
  for (long long j = 0; j < n; ++j) dst[j] = (double)src[j];
}
inline void widen_row_norm(const float* __restrict__ src, double* __restrict__ dst, long long n, double dr, const double* dc,
                           int nan_to_num) {
  static const bool avx2 = __builtin_cpu_supports("avx2");
  if (avx2) return widen_row_norm_avx2(src, dst, n, dr, dc, nan_to_num);
  for (long long j = 0; j < n; ++j) dst[j] = norm_one((double)src[j], dr, dc[j], nan_to_num);
}

// ---- transposed widening: dst[(j0 + b) * ld + i0 + a] = src[a * pitch + b]  for a < na, b < nb
__attribute__((target("avx2"))) inline void transpose_widen_avx2(const float* __restrict__ src, long long pitch, long long na, long long nb,
                                                                 double* __restrict__ dst, long long ld) {
  long long b = 0;
  for (; b + 8 <= nb; b += 8) {
    long long a = 0;
    for (; a + 8 <= na; a += 8) {
      const float* s = src + a * pitch + b;
      __m256 r0 = _mm256_loadu_ps(s), r1 = _mm256_loadu_ps(s + pitch), r2 = _mm256_loadu_ps(s + 2 * pitch),
             r3 = _mm256_loadu_ps(s + 3 * pitch), r4 = _mm256_loadu_ps(s + 4 * pitch), r5 = _mm256_loadu_ps(s + 5 * pitch),
             r6 = _mm256_loadu_ps(s + 6 * pitch), r7 = _mm256_loadu_ps(s + 7 * pitch);
      __m256 t0 = _mm256_unpacklo_ps(r0, r1), t1 = _mm256_unpackhi_ps(r0, r1), t2 = _mm256_unpacklo_ps(r2, r3),
             t3 = _mm256_unpackhi_ps(r2, r3), t4 = _mm256_unpacklo_ps(r4, r5), t5 = _mm256_unpackhi_ps(r4, r5),
             t6 = _mm256_unpacklo_ps(r6, r7), t7 = _mm256_unpackhi_ps(r6, r7);
      r0 = _mm256_shuffle_ps(t0, t2, 0x44); r1 = _mm256_shuffle_ps(t0, t2, 0xEE);
      r2 = _mm256_shuffle_ps(t1, t3, 0x44); r3 = _mm256_shuffle_ps(t1, t3, 0xEE);
      r4 = _mm256_shuffle_ps(t4, t6, 0x44); r5 = _mm256_shuffle_ps(t4, t6, 0xEE);
      r6 = _mm256_shuffle_ps(t5, t7, 0x44); r7 = _mm256_shuffle_ps(t5, t7, 0xEE);
      __m256 c[8];
      c[0] = _mm256_permute2f128_ps(r0, r4, 0x20); c[1] = _mm256_permute2f128_ps(r1, r5, 0x20);
      c[2] = _mm256_permute2f128_ps(r2, r6, 0x20); c[3] = _mm256_permute2f128_ps(r3, r7, 0x20);
      c[4] = _mm256_permute2f128_ps(r0, r4, 0x31); c[5] = _mm256_permute2f128_ps(r1, r5, 0x31);
      c[6] = _mm256_permute2f128_ps(r2, r6, 0x31); c[7] = _mm256_permute2f128_ps(r3, r7, 0x31);
      for (int k = 0; k < 8; ++k) {
        double* d = dst + (b + k) * ld + a;
        const __m256d lo = _mm256_cvtps_pd(_mm256_castps256_ps128(c[k])), hi = _mm256_cvtps_pd(_mm256_extractf128_ps(c[k], 1));
        if ((reinterpret_cast<uintptr_t>(d) & 31) == 0) { _mm256_stream_pd(d, lo); _mm256_stream_pd(d + 4, hi); }
        else { _mm256_storeu_pd(d, lo); _mm256_storeu_pd(d + 4, hi); }
      }
    }
    for (; a < na; ++a)
      for (int k = 0; k < 8; ++k) dst[(b + k) * ld + a] = (double)src[a * pitch + b + k];
  }
  for (; b < nb; ++b)
    for (long long a = 0; a < na; ++a) dst[b * ld + a] = (double)src[a * pitch + b];
}

inline void transpose_widen(const float* __restrict__ src, long long pitch, long long na, long long nb, double* __restrict__ dst,
                            long long ld) {
  static const bool avx2 = __builtin_cpu_supports("avx2");
  if (avx2) return transpose_widen_avx2(src, pitch, na, nb, dst, ld);
  for (long long b = 0; b < nb; ++b)
    for (long long a = 0; a < na; ++a) dst[b * ld + a] = (double)src[a * pitch + b];
}

// ---- the same two primitives for u16 sources (matrices whose entries are integers < 65 536 travel as 2 bytes)
__attribute__((target("avx2"))) inline void widen_row_u16_avx2(const uint16_t* __restrict__ src, double* __restrict__ dst, long long n) {
  long long j = 0;
  for (; j < n && (reinterpret_cast<uintptr_t>(dst + j) & 31); ++j) dst[j] = (double)src[j];
  for (; j + 8 <= n; j += 8) {
    const __m256i v = _mm256_cvtepu16_epi32(_mm_loadu_si128(reinterpret_cast<const __m128i*>(src + j)));
    _mm256_stream_pd(dst + j, _mm256_cvtepi32_pd(_mm256_castsi256_si128(v)));
    _mm256_stream_pd(dst + j + 4, _mm256_cvtepi32_pd(_mm256_extracti128_si256(v, 1)));
  }
  for (; j < n; ++j) dst[j] = (double)src[j];
}
inline void widen_row(const uint16_t* __restrict__ src, double* __restrict__ dst, long long n) {
  static const bool avx2 = __builtin_cpu_supports("avx2");
  if (avx2) return widen_row_u16_avx2(src, dst, n);
  for (long long j = 0; j < n; ++j) dst[j] = (double)src[j];
}
__attribute__((target("avx2"))) inline void transpose_widen_u16_avx2(const uint16_t* __restrict__ src, long long pitch, long long na,
                                                                     long long nb, double* __restrict__ dst, long long ld) {
  long long b = 0;
  for (; b + 8 <= nb; b += 8) {
    long long a = 0;
    for (; a + 8 <= na; a += 8) {
      const uint16_t* s = src + a * pitch + b;
      __m128i r[8];
      for (int k = 0; k < 8; ++k) r[k] = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + k * pitch));
      // 8 x 8 transpose of 16-bit lanes
      const __m128i t0 = _mm_unpacklo_epi16(r[0], r[1]), t1 = _mm_unpackhi_epi16(r[0], r[1]), t2 = _mm_unpacklo_epi16(r[2], r[3]),
                    t3 = _mm_unpackhi_epi16(r[2], r[3]), t4 = _mm_unpacklo_epi16(r[4], r[5]), t5 = _mm_unpackhi_epi16(r[4], r[5]),
                    t6 = _mm_unpacklo_epi16(r[6], r[7]), t7 = _mm_unpackhi_epi16(r[6], r[7]);
      const __m128i u0 = _mm_unpacklo_epi32(t0, t2), u1 = _mm_unpackhi_epi32(t0, t2), u2 = _mm_unpacklo_epi32(t1, t3),
                    u3 = _mm_unpackhi_epi32(t1, t3), u4 = _mm_unpacklo_epi32(t4, t6), u5 = _mm_unpackhi_epi32(t4, t6),
                    u6 = _mm_unpacklo_epi32(t5, t7), u7 = _mm_unpackhi_epi32(t5, t7);
      __m128i c[8];
      c[0] = _mm_unpacklo_epi64(u0, u4); c[1] = _mm_unpackhi_epi64(u0, u4); c[2] = _mm_unpacklo_epi64(u1, u5);
      c[3] = _mm_unpackhi_epi64(u1, u5); c[4] = _mm_unpacklo_epi64(u2, u6); c[5] = _mm_unpackhi_epi64(u2, u6);
      c[6] = _mm_unpacklo_epi64(u3, u7); c[7] = _mm_unpackhi_epi64(u3, u7);
      for (int k = 0; k < 8; ++k) {
        double* d = dst + (b + k) * ld + a;
        const __m256i v = _mm256_cvtepu16_epi32(c[k]);
        const __m256d lo = _mm256_cvtepi32_pd(_mm256_castsi256_si128(v)), hi = _mm256_cvtepi32_pd(_mm256_extracti128_si256(v, 1));
        if ((reinterpret_cast<uintptr_t>(d) & 31) == 0) { _mm256_stream_pd(d, lo); _mm256_stream_pd(d + 4, hi); }
        else { _mm256_storeu_pd(d, lo); _mm256_storeu_pd(d + 4, hi); }
      }
    }
    for (; a < na; ++a)
      for (int k = 0; k < 8; ++k) dst[(b + k) * ld + a] = (double)src[a * pitch + b + k];
  }
  for (; b < nb; ++b)
    for (long long a = 0; a < na; ++a) dst[b * ld + a] = (double)src[a * pitch + b];
}
inline void transpose_widen(const uint16_t* __restrict__ src, long long pitch, long long na, long long nb, double* __restrict__ dst,
                            long long ld) {
  static const bool avx2 = __builtin_cpu_supports("avx2");
  if (avx2) return transpose_widen_u16_avx2(src, pitch, na, nb, dst, ld);
  for (long long b = 0; b < nb; ++b)
    for (long long a = 0; a < na; ++a) dst[b * ld + a] = (double)src[a * pitch + b];
}

// Row bands of the upper triangle, of about `area` elements each: rows [start[c], start[c+1]), columns [start[c], n);
// row counts are multiples of 8.  Shared by the host delivery and the device-side u16 packer.
inline std::vector<long long> tri_bands(long long n) {
  long long area = 2LL << 20;
  if (const char* e = getenv("GRAKEL_B200_BAND_MB")) area = std::max(1, atoi(e)) * (1LL << 18);
  std::vector<long long> start;
  for (long long r = 0; r < n;) {
    start.push_back(r);
    long long nr = std::max<long long>(8, area / (n - r) / 8 * 8);
    if (r + nr + 8 > n) nr = n - r;  // no sliver at the end
    r += nr;
  }
  start.push_back(n);
  return start;
}

constexpr int DELIVER_SLOTS = 4;

// where the bands come from: the device (cudaMemcpy2DAsync + one event per staging slot) or, for the
// host-only self test of the band / transpose logic (gk_selftest_deliver), another host matrix
struct DeviceCopier {
  gk_handle* h;
  // pinned staging ring, allocated (and thereby first touched) while the calling thread sits on the GPU's NUMA node
  char* stage(size_t bytes) {
    if (bytes <= h->h_stage.cap) return h->h_stage.as<char>();
    cpu_set_t old_set, node_set;
    bool moved = false;
    std::vector<int> cpus = HostPool::local_node_cpus();
    if (!cpus.empty() && pthread_getaffinity_np(pthread_self(), sizeof(old_set), &old_set) == 0) {
      CPU_ZERO(&node_set);
      for (int c : cpus) CPU_SET(c, &node_set);
      moved = pthread_setaffinity_np(pthread_self(), sizeof(node_set), &node_set) == 0;
    }
    const int rc = h->h_stage.ensure(bytes);
    if (rc == GK_OK) memset(h->h_stage.p, 0, 4096);
    if (moved) pthread_setaffinity_np(pthread_self(), sizeof(old_set), &old_set);
    return rc == GK_OK ? h->h_stage.as<char>() : nullptr;
  }
  bool copy(int slot, void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height) {
    if (cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, cudaMemcpyDeviceToHost, h->stream) != cudaSuccess) return false;
    return cudaEventRecord(h->ev_stage[slot], h->stream) == cudaSuccess;
  }
  bool wait(int slot) { return cudaEventSynchronize(h->ev_stage[slot]) == cudaSuccess; }
};
struct HostCopier {
  std::vector<char> buf;
  char* stage(size_t bytes) { buf.resize(bytes); return buf.data(); }
  bool copy(int, void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height) {
    for (size_t r = 0; r < height; ++r) memcpy((char*)dst + r * dpitch, (const char*)src + r * spitch, width);
    return true;
  }
  bool wait(int) { return true; }
};

}  // namespace

// Row block [0, rows) x [0, cols) of an fp32 device matrix -> float64 host rows; optional normalisation with host
// diagonals (drow[r], dcol[c]).  Bands of ~8 MB through the staging ring.
template <class Copier>
static int deliver_rows(Copier& cp, const float* d_src, long long d_ld, long long rows, long long cols, double* dst,
                        long long ld, const double* drow, const double* dcol, int nan_to_num) {
  const long long band_rows = std::max<long long>(1, std::min<long long>(rows, (8LL << 20) / std::max<long long>(cols * 4, 1)));
  const size_t slot_bytes = ((size_t)band_rows * cols * 4 + 255) / 256 * 256;
  char* stage = cp.stage(slot_bytes * DELIVER_SLOTS);
  if (!stage) return GK_ERR_CUDA;
  const long long n_bands = (rows + band_rows - 1) / band_rows;
  HostPool& pool = host_pool();
  std::atomic<long long> ready(0);
  std::vector<std::atomic<int>> done(n_bands);
  for (auto& d : done) d.store(0);
  auto enqueue = [&](long long c) -> int {
    const long long r0 = c * band_rows, nr = std::min(band_rows, rows - r0);
    if (!cp.copy((int)(c % DELIVER_SLOTS), stage + (size_t)(c % DELIVER_SLOTS) * slot_bytes, (size_t)cols * 4, d_src + r0 * d_ld,
                 (size_t)d_ld * 4, (size_t)cols * 4, (size_t)nr))
      return fail(GK_ERR_CUDA, "deliver_rows: D2H copy could not be queued");
    return GK_OK;
  };
  // tasks: runs of 8 rows, handed out through one atomic counter (a delayed worker holds back 8 rows, not 1/nt of a band)
  const long long tasks_per_band = (band_rows + 7) / 8;
  std::atomic<long long> next(0);
  const std::function<void(int)> work = [&](int) {
    for (;;) {
      const long long t = next.fetch_add(1, std::memory_order_relaxed);
      const long long c = t / tasks_per_band;
      if (c >= n_bands) return;
      for (unsigned spins = 0; ready.load(std::memory_order_acquire) <= c; ++spins) {
        if (ready.load(std::memory_order_relaxed) < 0) return;  // aborted
        if ((spins & 1023u) == 1023u) sched_yield(); else _mm_pause();
      }
      const long long r0 = c * band_rows, nr = std::min(band_rows, rows - r0);
      const float* src = reinterpret_cast<const float*>(stage + (size_t)(c % DELIVER_SLOTS) * slot_bytes);
      const long long a = (t - c * tasks_per_band) * 8, b = std::min(nr, a + 8);
      for (long long r = a; r < b; ++r) {
        if (drow) widen_row_norm(src + r * cols, dst + (r0 + r) * ld, cols, drow[r0 + r], dcol, nan_to_num);
        else widen_row(src + r * cols, dst + (r0 + r) * ld, cols);
      }
      done[c].fetch_add(1, std::memory_order_release);
    }
  };
  int rc = GK_OK;
  for (long long c = 0; c < std::min<long long>(n_bands, DELIVER_SLOTS) && rc == GK_OK; ++c) rc = enqueue(c);
  if (rc != GK_OK) return rc;
  pool.start(work);
  for (long long c = 0; c < n_bands; ++c) {
    if (!cp.wait((int)(c % DELIVER_SLOTS))) { rc = fail(GK_ERR_CUDA, "deliver_rows: D2H copy failed"); break; }
    ready.store(c + 1, std::memory_order_release);
    if (c >= 1 && c - 1 + DELIVER_SLOTS < n_bands) {  // slot of band c-1 is free once every worker has left it
      while (done[c - 1].load(std::memory_order_acquire) < (int)tasks_per_band) _mm_pause();
      rc = enqueue(c - 1 + DELIVER_SLOTS);
      if (rc != GK_OK) break;
    }
  }
  if (rc != GK_OK) ready.store(-1, std::memory_order_release);
  pool.wait();
  _mm_sfence();
  return rc;
}

// Symmetric n x n fp32 device matrix -> full float64 host matrix; only the upper triangle is copied.
// T = float: `d_src` is the n x n matrix itself (pitch d_ld elements).  T = uint16_t: `d_src` is the band-packed upper
// triangle produced by pack_tri_u16 (band c at element offset sum_{b<c} rows_b * width_b, pitch = its own width).
template <class T, class Copier>
static int deliver_tri(Copier& cp, const T* d_src, long long d_ld, long long n, double* dst, long long ld) {
  const bool packed = sizeof(T) == 2;
  const std::vector<long long> start = tri_bands(n);
  const long long n_bands = (long long)start.size() - 1;
  size_t slot_bytes = 0;
  for (long long c = 0; c < n_bands; ++c)
    slot_bytes = std::max(slot_bytes, (size_t)(start[c + 1] - start[c]) * (size_t)(n - start[c]) * sizeof(T));
  slot_bytes = (slot_bytes + 255) / 256 * 256;
  char* stage = cp.stage(slot_bytes * DELIVER_SLOTS);
  if (!stage) return GK_ERR_CUDA;
  HostPool& pool = host_pool();
  std::atomic<long long> ready(0);
  std::vector<std::atomic<int>> done(n_bands);
  for (auto& d : done) d.store(0);
  std::vector<long long> packed_off(n_bands + 1, 0);
  for (long long c = 0; c < n_bands; ++c) packed_off[c + 1] = packed_off[c] + (start[c + 1] - start[c]) * (n - start[c]);
  auto enqueue = [&](long long c) -> int {
    const long long r0 = start[c], nr = start[c + 1] - r0, w = n - r0;
    const T* src = packed ? d_src + packed_off[c] : d_src + r0 * d_ld + r0;
    const size_t spitch = (size_t)(packed ? w : d_ld) * sizeof(T);
    if (!cp.copy((int)(c % DELIVER_SLOTS), stage + (size_t)(c % DELIVER_SLOTS) * slot_bytes, (size_t)w * sizeof(T), src, spitch,
                 (size_t)w * sizeof(T), (size_t)nr))
      return fail(GK_ERR_CUDA, "deliver_tri: D2H copy could not be queued");
    return GK_OK;
  };
  // tasks: 64-column strips of a band (straight part + its mirrored block), handed out through one atomic counter
  std::vector<long long> task0(n_bands + 1, 0);
  for (long long c = 0; c < n_bands; ++c) task0[c + 1] = task0[c] + (n - start[c] + 63) / 64;
  std::atomic<long long> next(0);
  const std::function<void(int)> work = [&](int) {
    long long c = 0;
    for (;;) {
      const long long t = next.fetch_add(1, std::memory_order_relaxed);
      if (t >= task0[n_bands]) return;
      while (t >= task0[c + 1]) ++c;
      for (unsigned spins = 0; ready.load(std::memory_order_acquire) <= c; ++spins) {
        if (ready.load(std::memory_order_relaxed) < 0) return;
        if ((spins & 1023u) == 1023u) sched_yield(); else _mm_pause();
      }
      const long long r0 = start[c], r1 = start[c + 1], nr = r1 - r0, w = n - r0;
      const T* src = reinterpret_cast<const T*>(stage + (size_t)(c % DELIVER_SLOTS) * slot_bytes);
      const long long ca = (t - task0[c]) * 64, cb = std::min(w, ca + 64);  // columns of the band, relative to r0
      for (long long r = 0; r < nr; ++r) widen_row(src + r * w + ca, dst + (r0 + r) * ld + r0 + ca, cb - ca);
      // mirrored part: band columns >= nr are rows r1.. of the destination, columns r0..r1
      const long long ma = std::max(ca, nr);
      if (cb > ma) transpose_widen(src + ma, w, nr, cb - ma, dst + (r0 + ma) * ld + r0, ld);
      done[c].fetch_add(1, std::memory_order_release);
    }
  };
  int rc = GK_OK;
  for (long long c = 0; c < std::min<long long>(n_bands, DELIVER_SLOTS) && rc == GK_OK; ++c) rc = enqueue(c);
  if (rc != GK_OK) return rc;
  pool.start(work);
  for (long long c = 0; c < n_bands; ++c) {
    if (!cp.wait((int)(c % DELIVER_SLOTS))) { rc = fail(GK_ERR_CUDA, "deliver_tri: D2H copy failed"); break; }
    ready.store(c + 1, std::memory_order_release);
    if (c >= 1 && c - 1 + DELIVER_SLOTS < n_bands) {
      while (done[c - 1].load(std::memory_order_acquire) < (int)(task0[c] - task0[c - 1])) _mm_pause();
      rc = enqueue(c - 1 + DELIVER_SLOTS);
      if (rc != GK_OK) break;
    }
  }
  if (rc != GK_OK) ready.store(-1, std::memory_order_release);
  pool.wait();
  _mm_sfence();
  return rc;
}

// ---- result buffers for float64 matrices: anonymous mappings with transparent huge pages requested, recycled
// through a small pool so that a loop of fit_transform calls does not fault 800 MB of fresh pages in every time
namespace {
struct HostBlock { void* p; size_t bytes; };
std::mutex g_hostpool_m;
std::vector<HostBlock> g_hostpool;
size_t g_hostpool_bytes = 0;
size_t hostpool_limit() {
  static const size_t lim = [] {
    size_t mb = 2048;
    if (const char* e = getenv("GRAKEL_B200_HOST_POOL_MB")) mb = (size_t)std::max(0, atoi(e));
    return mb << 20;
  }();
  return lim;
}
}  // namespace

extern "C" {

// Host-only self test of the delivery code (no GPU): `src` is an n x n fp32 matrix on the host (symmetric for
// mode 0).  mode 0: deliver_tri; mode 1: deliver_rows; mode 2: deliver_rows with normalisation by diag.
int gk_selftest_deliver(int32_t mode, int64_t rows, int64_t cols, const float* src, const double* diag, int32_t nan_to_num,
                        double* dst) {
  if (!src || !dst || rows <= 0 || cols <= 0) return fail(GK_ERR_ARG, "gk_selftest_deliver: bad arguments");
  HostCopier cp;
  if (mode == 0 || mode == 3) {
    if (rows != cols) return fail(GK_ERR_ARG, "gk_selftest_deliver: modes 0 and 3 need a square matrix");
    if (mode == 0) return deliver_tri<float>(cp, src, cols, rows, dst, cols);
    // mode 3: what pack_tri_u16 does on the device, on the host -- then the u16 delivery
    const std::vector<long long> start = tri_bands(rows);
    std::vector<uint16_t> packed;
    for (size_t c = 0; c + 1 < start.size(); ++c)
      for (long long r = start[c]; r < start[c + 1]; ++r)
        for (long long j = start[c]; j < cols; ++j) packed.push_back((uint16_t)src[r * cols + j]);
    return deliver_tri<uint16_t>(cp, packed.data(), 0, rows, dst, cols);
  }
  return deliver_rows(cp, src, cols, rows, cols, dst, cols, mode == 2 ? diag : nullptr, mode == 2 ? diag : nullptr, nan_to_num);
}

int gk_host_alloc(int64_t bytes, void** out) {
  if (!out || bytes <= 0) return fail(GK_ERR_ARG, "gk_host_alloc: bad arguments");
  const size_t want = ((size_t)bytes + (2u << 20) - 1) / (2u << 20) * (2u << 20);
  {
    std::lock_guard<std::mutex> l(g_hostpool_m);
    for (size_t i = 0; i < g_hostpool.size(); ++i)
      if (g_hostpool[i].bytes == want) {
        *out = g_hostpool[i].p;
        g_hostpool_bytes -= want;
        g_hostpool.erase(g_hostpool.begin() + i);
        return GK_OK;
      }
  }
  void* p = mmap(nullptr, want, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
  if (p == MAP_FAILED) return fail(GK_ERR_CUDA, "gk_host_alloc: mmap failed");
  madvise(p, want, MADV_HUGEPAGE);  // best effort
  *out = p;
  return GK_OK;
}

int gk_host_free(void* p, int64_t bytes) {
  if (!p || bytes <= 0) return GK_OK;
  const size_t want = ((size_t)bytes + (2u << 20) - 1) / (2u << 20) * (2u << 20);
  {
    std::lock_guard<std::mutex> l(g_hostpool_m);
    if (g_hostpool_bytes + want <= hostpool_limit()) {
      g_hostpool.push_back({p, want});
      g_hostpool_bytes += want;
      return GK_OK;
    }
  }
  munmap(p, want);
  return GK_OK;
}

}  // extern "C"
