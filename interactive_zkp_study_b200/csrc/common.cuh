// Shared host-side plumbing for libzkp_b200: error reporting, the per-process device context
// (one device per process, one process per GPU), a grow-only workspace arena.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace zkp {

struct CudaError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

inline void cuda_check(cudaError_t e, const char* what, const char* file, int line) {
  if (e != cudaSuccess) {
    char buf[512];
    snprintf(buf, sizeof buf, "%s failed at %s:%d: %s", what, file, line, cudaGetErrorString(e));
    throw CudaError(buf);
  }
}
#define CUDA_CHECK(x) ::zkp::cuda_check((x), #x, __FILE__, __LINE__)
#define CUDA_CHECK_LAUNCH() ::zkp::cuda_check(cudaGetLastError(), "kernel launch", __FILE__, __LINE__)

// Every device allocation of the library goes through these two.  With ZKP_B200_GUARD=1 in the environment each
// block gets a 512-byte canary zone on both sides (filled with 0xA5, checked when the block is freed and by
// zkp_debug_check_guards): compute-sanitizer is closed on the B200 pool, so this is how an out-of-bounds WRITE
// next to any workspace, table or scalar vector is caught (capi_core.cu).
void* dev_alloc(size_t bytes);
void dev_free(void* p);

// A device buffer that only ever grows; reused across calls so steady-state calls allocate nothing.
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  void reserve(size_t bytes) {
    if (bytes <= cap) return;
    if (p) dev_free(p);
    p = nullptr;
    cap = 0;
    p = dev_alloc(bytes);
    cap = bytes;
  }
  template <class T>
  T* as() const { return reinterpret_cast<T*>(p); }
  void release() {
    if (p) dev_free(p);
    p = nullptr;
    cap = 0;
  }
  // Handle-owned buffers come from / go back to a small size-matched pool: cudaMalloc / cudaFree of
  // tens of MiB cost 0.1-1 ms each (cudaFree also synchronises the device), which would dominate a
  // proof that assembles a handful of scalar vectors.
  void reserve_pooled(size_t bytes);
  void recycle();
};

// A temporary device buffer owned by one scope: released when the scope ends, also when it ends by an
// exception (DevBuf itself is a plain handle that long-lived workspaces copy around, so it has no destructor).
struct ScopedDevBuf : DevBuf {
  ScopedDevBuf() = default;
  ScopedDevBuf(const ScopedDevBuf&) = delete;
  ScopedDevBuf& operator=(const ScopedDevBuf&) = delete;
  ~ScopedDevBuf() { release(); }
  // hands the allocation over to a long-lived owner
  DevBuf detach() {
    DevBuf out;
    out.p = p;
    out.cap = cap;
    p = nullptr;
    cap = 0;
    return out;
  }
};
// Same for buffers taken from the size-matched pool (reserve_pooled): back to the pool at scope end.
struct PooledDevBuf : DevBuf {
  PooledDevBuf() = default;
  PooledDevBuf(const PooledDevBuf&) = delete;
  PooledDevBuf& operator=(const PooledDevBuf&) = delete;
  ~PooledDevBuf() { recycle(); }
};

// [offset, offset + n) inside [0, limit) without wrapping: ctypes turns a negative Python int into
// 2^64 - k, and `offset + n > limit` would let that through.
inline bool range_ok(uint64_t offset, uint64_t n, uint64_t limit) { return offset <= limit && n <= limit - offset; }

struct Context {
  int device = -1;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;                 // second lane of the batched-MSM pipeline
  cudaStream_t stream_hi = nullptr;               // high-priority side lane of the part-streamed MSM (upload + sort)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  std::mutex mu;  // serialises entry points (Flask's dev server is threaded, ctypes drops the GIL)
  unsigned long long launches = 0;  // kernels launched by this library (bench.py "gpu_launches")
};

// Per-stage CUDA-event profile of the last MSM (enabled by zkp_msm_profile(1) or ZKP_B200_TRACE=1):
// device time between marks, kept for zkp_msm_last_profile and optionally printed.
struct StageProfile {
  std::vector<std::pair<std::string, float>> stages;  // (name, microseconds)
  float total_us = 0;
};
StageProfile& last_msm_profile();
extern int g_msm_profile_enabled;

struct StageTrace {
  bool on, print;
  cudaStream_t st;
  std::vector<std::pair<const char*, cudaEvent_t>> ev;
  // events are created once and reused: creating/destroying ~20 events per MSM costs more than
  // recording them, and the profile is taken inside bench.py's timed region
  static std::vector<cudaEvent_t>& pool() {
    static std::vector<cudaEvent_t> p;
    return p;
  }
  explicit StageTrace(cudaStream_t s, bool enabled = true) : st(s) {
    static const bool env = getenv("ZKP_B200_TRACE") && atoi(getenv("ZKP_B200_TRACE"));
    print = env && enabled;
    on = (env || g_msm_profile_enabled) && enabled;
    mark("start");
  }
  void mark(const char* name) {
    if (!on) return;
    // bench.py only needs the accumulate kernel and the total: when not printing, record just the
    // marks that bracket them (4 events per MSM instead of ~22 inside the timed region)
    if (!print && strcmp(name, "start") && strcmp(name, "tasks") && strcmp(name, "accumulate") &&
        strcmp(name, "horner+affine"))
      return;
    std::vector<cudaEvent_t>& p = pool();
    if (ev.size() >= p.size()) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      p.push_back(e);
    }
    cudaEvent_t e = p[ev.size()];
    cudaEventRecord(e, st);
    ev.emplace_back(name, e);
  }
  ~StageTrace() {
    if (!on || ev.empty()) return;
    cudaEventSynchronize(ev.back().second);
    StageProfile& prof = last_msm_profile();
    prof.stages.clear();
    for (size_t i = 1; i < ev.size(); i++) {
      float ms = 0;
      cudaEventElapsedTime(&ms, ev[i - 1].second, ev[i].second);
      prof.stages.emplace_back(ev[i].first, ms * 1e3f);
      if (print) fprintf(stderr, "[zkp trace] %-18s %8.1f us\n", ev[i].first, ms * 1e3);
    }
    float tot = 0;
    cudaEventElapsedTime(&tot, ev.front().second, ev.back().second);
    prof.total_us = tot * 1e3f;
    if (print) fprintf(stderr, "[zkp trace] %-18s %8.1f us\n", "TOTAL", tot * 1e3);
  }
};

Context& ctx();             // throws if zkp_init has not succeeded
bool ctx_ready();

inline unsigned int ceil_div(uint64_t a, uint64_t b) { return (unsigned int)((a + b - 1) / b); }

}  // namespace zkp
