// libzkp_b200: lifecycle, handles, timers, diagnostics.  See include/zkp_b200.h.
#include <cstdlib>
#include <cstring>
#include <unordered_map>
#include "ff.cuh"
#include "registry.cuh"

namespace zkp {

static Context g_ctx;
static bool g_ready = false;
static std::mutex g_init_mu;
// the pool is declared before the registry: at exit the registry goes first and its resources hand
// their buffers back to a pool that still exists
static std::vector<DevBuf> g_pool;
static size_t g_pool_bytes = 0;
static Registry g_registry;
static std::mutex g_err_mu;
static std::string g_last_error;
static cudaEvent_t g_ev0 = nullptr, g_ev1 = nullptr;

static StageProfile g_last_profile;
int g_msm_profile_enabled = 0;
StageProfile& last_msm_profile() { return g_last_profile; }

// ------------------------------------------------------------------ guarded allocation (ZKP_B200_GUARD=1)
static constexpr size_t GUARD_BYTES = 512;
static constexpr unsigned char GUARD_FILL = 0xA5;
static std::mutex g_guard_mu;
static std::unordered_map<void*, size_t> g_guarded;  // user pointer -> user bytes
static unsigned long long g_guard_violations = 0;

static bool guard_mode() {
  static const bool on = getenv("ZKP_B200_GUARD") && atoi(getenv("ZKP_B200_GUARD"));
  return on;
}

// number of canary bytes around `user` that no longer hold the fill pattern (device synchronised first)
static size_t guard_damage(void* user, size_t bytes) {
  unsigned char host[2 * GUARD_BYTES];
  char* base = static_cast<char*>(user) - GUARD_BYTES;
  cudaDeviceSynchronize();
  if (cudaMemcpy(host, base, GUARD_BYTES, cudaMemcpyDeviceToHost) != cudaSuccess) return 1;
  if (cudaMemcpy(host + GUARD_BYTES, base + GUARD_BYTES + bytes, GUARD_BYTES, cudaMemcpyDeviceToHost) != cudaSuccess) return 1;
  size_t bad = 0;
  for (size_t i = 0; i < 2 * GUARD_BYTES; i++) bad += host[i] != GUARD_FILL;
  return bad;
}

void* dev_alloc(size_t bytes) {
  void* p = nullptr;
  if (!guard_mode()) {
    CUDA_CHECK(cudaMalloc(&p, bytes));
    return p;
  }
  const size_t padded = (bytes + 15) & ~size_t(15);  // the trailing canary starts on a 16-byte boundary
  CUDA_CHECK(cudaMalloc(&p, padded + 2 * GUARD_BYTES));
  char* base = static_cast<char*>(p);
  CUDA_CHECK(cudaMemset(base, GUARD_FILL, GUARD_BYTES));
  CUDA_CHECK(cudaMemset(base + GUARD_BYTES + padded, GUARD_FILL, GUARD_BYTES));
  std::lock_guard<std::mutex> lk(g_guard_mu);
  g_guarded[base + GUARD_BYTES] = padded;
  return base + GUARD_BYTES;
}

void dev_free(void* p) {
  if (!p) return;
  if (!guard_mode()) {
    cudaFree(p);
    return;
  }
  size_t bytes = 0;
  {
    std::lock_guard<std::mutex> lk(g_guard_mu);
    auto it = g_guarded.find(p);
    if (it == g_guarded.end()) return;  // not ours (or freed twice): leave it
    bytes = it->second;
    g_guarded.erase(it);
  }
  size_t bad = guard_damage(p, bytes);
  if (bad) {
    g_guard_violations++;
    fprintf(stderr, "[zkp guard] %zu canary bytes overwritten around a %zu-byte device buffer\n", bad, bytes);
  }
  cudaFree(static_cast<char*>(p) - GUARD_BYTES);
}

// ------------------------------------------------------------------ device buffer pool
static constexpr size_t POOL_MAX_BYTES = size_t(8) << 30;

void DevBuf::reserve_pooled(size_t bytes) {
  if (bytes <= cap) return;
  if (p) recycle();
  int best = -1;
  for (int i = 0; i < (int)g_pool.size(); i++)
    if (g_pool[i].cap >= bytes && g_pool[i].cap <= 2 * bytes + 4096 && (best < 0 || g_pool[i].cap < g_pool[best].cap)) best = i;
  if (best >= 0) {
    p = g_pool[best].p;
    cap = g_pool[best].cap;
    g_pool_bytes -= cap;
    g_pool.erase(g_pool.begin() + best);
    return;
  }
  reserve(bytes);
}

void DevBuf::recycle() {
  if (!p) return;
  if (cap > (size_t(2) << 30)) {
    release();
    return;
  }
  // over budget: the OLDEST pooled buffers go (leftovers of an earlier phase), the one coming in stays -- it
  // belongs to the working set of whatever is running now.  (Dropping the newcomer instead made a prover that ran
  // after other work pay a cudaMalloc + cudaFree per vector: PLONK 2^20 47 ms instead of 40 inside bench.py.)
  while (!g_pool.empty() && g_pool_bytes + cap > POOL_MAX_BYTES) {
    g_pool_bytes -= g_pool.front().cap;
    g_pool.front().release();
    g_pool.erase(g_pool.begin());
  }
  DevBuf b;
  b.p = p;
  b.cap = cap;
  g_pool.push_back(b);
  g_pool_bytes += cap;
  p = nullptr;
  cap = 0;
}

Context& ctx() {
  if (!g_ready) throw std::runtime_error("zkp_init has not been called successfully");
  return g_ctx;
}
bool ctx_ready() { return g_ready; }
Registry& registry() { return g_registry; }
void set_last_error(const std::string& s) {
  std::lock_guard<std::mutex> lk(g_err_mu);
  g_last_error = s;
}

// ------------------------------------------------------------------ synthetic scalars
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__global__ void scalars_generate_kernel(uint64_t seed, uint64_t n, uint32_t* out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t v[8];
  bool ok = false;
  for (int t = 0; t < 16 && !ok; t++) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
      uint64_t w = mix64(seed + 64 * i + 4 * t + j);
      v[2 * j] = (uint32_t)w;
      v[2 * j + 1] = (uint32_t)(w >> 32);
    }
    v[7] &= 0x3fffffffu;  // 254 bits
    // v < r ?
    uint32_t borrow = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      uint64_t d = (uint64_t)v[k] - FrParams::MOD_(k) - borrow;
      borrow = (uint32_t)(d >> 63);
    }
    ok = borrow != 0;
  }
  if (!ok) {  // 2^254 < 2r: one subtraction suffices
    uint32_t borrow = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      uint64_t d = (uint64_t)v[k] - FrParams::MOD_(k) - borrow;
      v[k] = (uint32_t)d;
      borrow = (uint32_t)(d >> 63);
    }
  }
  uint4* o = reinterpret_cast<uint4*>(out + 8 * i);
  o[0] = make_uint4(v[0], v[1], v[2], v[3]);
  o[1] = make_uint4(v[4], v[5], v[6], v[7]);
}

}  // namespace zkp

using namespace zkp;

extern "C" {

int zkp_init(int device) {
  try {
    std::lock_guard<std::mutex> lk(g_init_mu);  // two threads may make their first call at the same time
    if (g_ready) return ZKP_OK;
    int count = 0;
    CUDA_CHECK(cudaGetDeviceCount(&count));
    if (count == 0) throw std::runtime_error("no CUDA device visible; libzkp_b200 has no CPU fallback");
    if (device < 0) {
      const char* lr = getenv("LOCAL_RANK");
      device = lr ? atoi(lr) : 0;
    }
    if (device >= count) throw std::runtime_error("device index out of range");
    cudaDeviceProp prop;
    CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    // the library carries one sm_100a cubin and no PTX: any other part (sm_103 included) has no kernel image
    if (prop.major != 10 || prop.minor != 0)
      throw std::runtime_error(std::string("libzkp_b200 is built for sm_100a (compute capability 10.0) only; device is ") +
                               prop.name + " (" + std::to_string(prop.major) + "." + std::to_string(prop.minor) + ")");
    CUDA_CHECK(cudaSetDevice(device));
    g_ctx.device = device;
    g_ctx.sm_count = prop.multiProcessorCount;
    CUDA_CHECK(cudaStreamCreateWithFlags(&g_ctx.stream, cudaStreamNonBlocking));
    CUDA_CHECK(cudaStreamCreateWithFlags(&g_ctx.stream2, cudaStreamNonBlocking));
    {
      int lo = 0, hi = 0;  // numerically lower = higher priority
      CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      CUDA_CHECK(cudaStreamCreateWithPriority(&g_ctx.stream_hi, cudaStreamNonBlocking, hi));
    }
    CUDA_CHECK(cudaEventCreateWithFlags(&g_ctx.ev_fork, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&g_ctx.ev_join, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreate(&g_ev0));
    CUDA_CHECK(cudaEventCreate(&g_ev1));
    g_ready = true;
    return ZKP_OK;
  } catch (const std::exception& e) {
    set_last_error(e.what());
    return ZKP_ERR_CUDA;
  }
}

int zkp_shutdown(void) {
  if (!g_ready) return ZKP_OK;
  std::lock_guard<std::mutex> lk(g_ctx.mu);
  cudaSetDevice(g_ctx.device);
  cudaStreamSynchronize(g_ctx.stream);
  for (auto& kv : g_registry.items) {
    kv.second->buf.release();
    for (DevBuf& b : kv.second->aux) b.release();
  }
  g_registry.items.clear();
  return ZKP_OK;  // the context (stream, engines' workspaces) stays usable; nothing else to tear down
}

const char* zkp_last_error(void) {
  static thread_local std::string copy;
  std::lock_guard<std::mutex> lk(g_err_mu);
  copy = g_last_error;
  return copy.c_str();
}

int zkp_device_info(char* name, int name_cap, int* sm_count, int* cc_major, int* cc_minor, int* sm_clock_khz) {
  return guarded([&](Context& c) {
    cudaDeviceProp prop;
    CUDA_CHECK(cudaGetDeviceProperties(&prop, c.device));
    if (name && name_cap > 0) {
      strncpy(name, prop.name, name_cap - 1);
      name[name_cap - 1] = 0;
    }
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (sm_clock_khz) {
      int khz = 0;
      CUDA_CHECK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, c.device));
      *sm_clock_khz = khz;
    }
  });
}

int zkp_device_pci_bus_id(char* out, int cap) {
  return guarded([&](Context& c) {
    if (!out || cap < 16) throw InvalidArgument("zkp_device_pci_bus_id: need a buffer of at least 16 bytes");
    CUDA_CHECK(cudaDeviceGetPCIBusId(out, cap, c.device));
  });
}

int zkp_debug_check_guards(uint64_t* live_buffers, uint64_t* violations) {
  return guarded([&](Context&) {
    if (!guard_mode()) throw InvalidArgument("zkp_debug_check_guards: start the process with ZKP_B200_GUARD=1");
    std::vector<std::pair<void*, size_t>> live;
    {
      std::lock_guard<std::mutex> lk(g_guard_mu);
      live.assign(g_guarded.begin(), g_guarded.end());
    }
    unsigned long long bad_buffers = 0;
    for (auto& kv : live) {
      size_t bad = guard_damage(kv.first, kv.second);
      if (bad) {
        bad_buffers++;
        fprintf(stderr, "[zkp guard] %zu canary bytes overwritten around a live %zu-byte device buffer\n", bad, kv.second);
      }
    }
    if (live_buffers) *live_buffers = live.size();
    if (violations) *violations = g_guard_violations + bad_buffers;
  });
}

int zkp_device_mem_info(uint64_t* free_bytes, uint64_t* total_bytes) {
  return guarded([&](Context&) {
    size_t f = 0, t = 0;
    CUDA_CHECK(cudaMemGetInfo(&f, &t));
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
  });
}

uint64_t zkp_launch_count(void) { return g_ready ? g_ctx.launches : 0; }

int zkp_timer_start(void) {
  return guarded([&](Context& c) { CUDA_CHECK(cudaEventRecord(g_ev0, c.stream)); });
}
int zkp_timer_stop(float* elapsed_ms) {
  return guarded([&](Context& c) {
    CUDA_CHECK(cudaEventRecord(g_ev1, c.stream));
    CUDA_CHECK(cudaEventSynchronize(g_ev1));
    CUDA_CHECK(cudaEventElapsedTime(elapsed_ms, g_ev0, g_ev1));
  });
}
int zkp_sync(void) {
  return guarded([&](Context& c) { CUDA_CHECK(cudaStreamSynchronize(c.stream)); });
}

int zkp_free(uint64_t handle) {
  return guarded([&](Context& c) {
    auto it = registry().items.find(handle);
    if (it == registry().items.end()) throw BadHandle("zkp_free: unknown handle");
    CUDA_CHECK(cudaStreamSynchronize(c.stream));
    it->second->buf.recycle();
    for (DevBuf& b : it->second->aux) b.recycle();
    registry().items.erase(it);
  });
}

int zkp_scalars_load(const uint8_t* scalars, uint64_t n, uint64_t* handle) {
  return guarded([&](Context& c) {
    if (!handle || (n && !scalars)) throw InvalidArgument("zkp_scalars_load: null argument");
    auto r = std::make_unique<Resource>();
    r->kind = HandleKind::Scalars;
    r->n = n;
    r->buf.reserve_pooled(n ? n * 32 : 32);
    if (n) CUDA_CHECK(cudaMemcpyAsync(r->buf.p, scalars, n * 32, cudaMemcpyHostToDevice, c.stream));
    CUDA_CHECK(cudaStreamSynchronize(c.stream));
    *handle = registry().put(std::move(r));
  });
}

int zkp_scalars_generate(uint64_t seed, uint64_t n, uint64_t* handle) {
  return guarded([&](Context& c) {
    if (!handle) throw InvalidArgument("zkp_scalars_generate: null handle");
    auto r = std::make_unique<Resource>();
    r->kind = HandleKind::Scalars;
    r->n = n;
    r->buf.reserve_pooled(n ? n * 32 : 32);
    if (n) {
      scalars_generate_kernel<<<ceil_div(n, 256), 256, 0, c.stream>>>(seed, n, r->buf.as<uint32_t>());
      CUDA_CHECK_LAUNCH();
      c.launches++;
    }
    CUDA_CHECK(cudaStreamSynchronize(c.stream));
    *handle = registry().put(std::move(r));
  });
}

int zkp_scalars_download(uint64_t scalars, uint64_t offset, uint64_t n, uint8_t* out) {
  return guarded([&](Context& c) {
    Resource* r = need(scalars, HandleKind::Scalars, "zkp_scalars_download");
    if (!range_ok(offset, n, r->n)) throw InvalidArgument("zkp_scalars_download: range out of bounds");
    CUDA_CHECK(cudaMemcpyAsync(out, r->buf.as<uint8_t>() + offset * 32, n * 32, cudaMemcpyDeviceToHost, c.stream));
    CUDA_CHECK(cudaStreamSynchronize(c.stream));
  });
}

int zkp_msm_profile(int enable) {
  g_msm_profile_enabled = enable ? 1 : 0;
  return ZKP_OK;
}

// Sums, over the stages of the last MSM, the device time of every stage whose name contains `stage`
// (NULL or "" = whole MSM).  Stage names: "digits+scan", "scatter", "tasks", "accumulate", "fold",
// "ws wide", "ws2", "horner+affine".
int zkp_msm_last_profile(const char* stage, float* out_us) {
  if (!out_us) return ZKP_ERR_INVALID_ARGUMENT;
  const StageProfile& p = last_msm_profile();
  if (!stage || !*stage) {
    *out_us = p.total_us;
    return ZKP_OK;
  }
  float acc = 0;
  for (auto& kv : p.stages)
    if (kv.first.find(stage) != std::string::npos) acc += kv.second;
  *out_us = acc;
  return ZKP_OK;
}

int zkp_pinned_alloc(uint64_t bytes, void** out) {
  return guarded([&](Context&) {
    if (!out) throw InvalidArgument("zkp_pinned_alloc: null out");
    CUDA_CHECK(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
  });
}
int zkp_pinned_free(void* p) {
  return guarded([&](Context&) { CUDA_CHECK(cudaFreeHost(p)); });
}

}  // extern "C"
