// libzkp_b200: lifecycle, handles, timers, diagnostics.  See include/zkp_b200.h.
#include <cstdlib>
#include <cstring>
#include "ec.cuh"
#include "ec_quad.cuh"
#include "registry.cuh"

namespace zkp {

static Context g_ctx;
static bool g_ready = false;
static Registry g_registry;
static std::mutex g_err_mu;
static std::string g_last_error;
static cudaEvent_t g_ev0 = nullptr, g_ev1 = nullptr;

static StageProfile g_last_profile;
int g_msm_profile_enabled = 0;
StageProfile& last_msm_profile() { return g_last_profile; }

// ------------------------------------------------------------------ device buffer pool
static std::vector<DevBuf> g_pool;
static size_t g_pool_bytes = 0;
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
  if (g_pool_bytes + cap > POOL_MAX_BYTES || cap > (size_t(2) << 30)) {
    release();
    return;
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

// ------------------------------------------------------------------ integer-MAD microbenchmark
// 8 independent dependent-chains per thread (the multiplicand of each MAD is the previous result,
// so nothing can be hoisted), 256 threads x 8 blocks/SM: the FMA-pipe issue rate is the only limit.
// VARIANT 0: mad.wide.u32 (IMAD.WIDE.U32, 32x32+64 -> 64: the roofline unit "limb-MAC")
//         1: mad.lo.u32   (IMAD)
//         2: mad.hi.u32   (IMAD.HI)
//         3: Fp Montgomery multiplication chains (136 limb-MACs each): the practical ceiling
static constexpr int PEAK_UNROLL = 16;

// One thread, dependent chains: the latency a lone thread pays per operation (what bounds the MSM's
// reduction tail).  mode 0/1/2: 1/2/4 independent Fp product chains per iteration (time per iteration);
// 3: XYZZ add, inlined products; 4: XYZZ add, out-of-line products; 5: XYZZ mixed add; 6: XYZZ double.
template <class F>
__device__ __noinline__ uint32_t quad_probe(int mode, int iters) {
  const int ql = threadIdx.x & 3;
  XYZZ<F> acc, q;
  for (int k = 0; k < 8; k++) {
    uint32_t m = k == 7 ? 0x0fffffffu : 0xffffffffu;
    acc.x.v[k] = (0x9e3779b9u * (k + 1)) & m; acc.y.v[k] = (0x9e3779b9u * (k + 9)) & m;
    acc.zz.v[k] = (0x9e3779b9u * (k + 17)) & m; acc.zzz.v[k] = (0x9e3779b9u * (k + 25)) & m;
    q.x.v[k] = (0x85ebca6bu * (k + 3)) & m; q.y.v[k] = (0x85ebca6bu * (k + 11)) & m;
    q.zz.v[k] = (0x85ebca6bu * (k + 19)) & m; q.zzz.v[k] = (0x85ebca6bu * (k + 27)) & m;
  }
  for (int it = 0; it < iters; it++) {
    if (mode == 0) acc = Quad<F>::add(acc, q, ql);
    else acc = Quad<F>::dbl(acc, ql);
  }
  return acc.x.v[0] ^ acc.zzz.v[0];
}

__global__ void latency_probe_kernel(int mode, int iters, uint32_t* out) {
  if (mode >= 7) {  // one warp, quad operations: 7 add (inlined products), 8 add (out-of-line), 9 double (out-of-line)
    uint32_t s = mode == 7 ? quad_probe<Fp>(0, iters) : quad_probe<FpC>(mode == 8 ? 0 : 1, iters);
    if (threadIdx.x == 0) out[0] = s;
    return;
  }
  if (threadIdx.x != 0) return;
  Fp x[4], y;
  for (int j = 0; j < 4; j++)
#pragma unroll
    for (int k = 0; k < 8; k++) x[j].v[k] = (0x9e3779b9u * (k + 1 + 8 * j)) & (k == 7 ? 0x0fffffffu : 0xffffffffu);
#pragma unroll
  for (int k = 0; k < 8; k++) y.v[k] = (0x85ebca6bu * (k + 3)) & (k == 7 ? 0x0fffffffu : 0xffffffffu);
  uint32_t s = 0;
  if (mode <= 2) {
    for (int it = 0; it < iters; it++) {
      x[0] = x[0] * y;
      if (mode >= 1) x[1] = x[1] * y;
      if (mode >= 2) { x[2] = x[2] * y; x[3] = x[3] * y; }
    }
    for (int j = 0; j < 4; j++) s ^= x[j].v[0];
  } else if (mode == 3 || mode == 5 || mode == 6) {
    XYZZ<Fp> acc, q;
    acc.x = x[0]; acc.y = x[1]; acc.zz = x[2]; acc.zzz = x[3];
    q.x = y; q.y = x[1] * y; q.zz = x[2] * y; q.zzz = x[3] * y;
    Affine<Fp> qa;
    qa.x = q.x; qa.y = q.y;
    for (int it = 0; it < iters; it++) {
      if (mode == 3) acc.add(q);
      else if (mode == 5) acc.madd(qa);
      else acc = acc.dbl();
    }
    s = acc.x.v[0] ^ acc.zzz.v[0];
  } else {
    XYZZ<FpC> acc, q;
    for (int k = 0; k < 8; k++) {
      acc.x.v[k] = x[0].v[k]; acc.y.v[k] = x[1].v[k]; acc.zz.v[k] = x[2].v[k]; acc.zzz.v[k] = x[3].v[k];
      q.x.v[k] = y.v[k]; q.y.v[k] = x[1].v[k] ^ 5; q.zz.v[k] = x[2].v[k] ^ 9; q.zzz.v[k] = x[3].v[k] ^ 3;
    }
    for (int it = 0; it < iters; it++) acc.add(q);
    s = acc.x.v[0] ^ acc.zzz.v[0];
  }
  out[0] = s;
}
template <int VARIANT>
__global__ void __launch_bounds__(256) imad_peak_kernel(uint32_t* out, int iters, uint32_t m0) {
  uint32_t seed = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  uint32_t b = m0 | 1u;
  if (VARIANT == 3) {
    Fp x, y;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      x.v[k] = seed + k * 0x9e3779b9u;
      y.v[k] = (seed ^ 0x5bd1e995u) + k * 0x85ebca6bu;
    }
    x.v[7] &= 0x0fffffffu;
    y.v[7] &= 0x0fffffffu;
    for (int it = 0; it < iters; it++) {
      x = x * y;
      y = y * x;
    }
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= x.v[k] ^ y.v[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    return;
  }
  unsigned long long acc[8];
#pragma unroll
  for (int k = 0; k < 8; k++) acc[k] = ((unsigned long long)(seed + k) << 32) | (seed * (k + 3));
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < PEAK_UNROLL; u++) {
#pragma unroll
      for (int k = 0; k < 8; k++) {
        if (VARIANT == 0) {
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"((uint32_t)acc[k]), "r"(b));
        } else if (VARIANT == 1) {
          uint32_t x = (uint32_t)acc[k];
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(seed));
          acc[k] = x;
        } else {
          uint32_t x = (uint32_t)acc[k];
          asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(seed));
          acc[k] = x;
        }
      }
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) s ^= (uint32_t)acc[k] ^ (uint32_t)(acc[k] >> 32);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ------------------------------------------------------------------ debug kernels (tests only)
template <class FE>
__global__ void dbg_field_op_kernel(int op, const FE* a, const FE* b, uint64_t n, FE* out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  FE x = a[i].to_mont(), y = b ? b[i].to_mont() : FE::zero(), r;
  switch (op) {
    case 0: r = x + y; break;
    case 1: r = x - y; break;
    case 2: r = x * y; break;
    case 3: r = x.inv(); break;
    case 5: r = x.inv_fermat(); break;
    default: r = x.sqr(); break;
  }
  out[i] = r.from_mont();
}

template <class F>
__global__ void dbg_point_add_kernel(const Affine<F>* a, const Affine<F>* b, uint64_t n, Affine<F>* out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine<F> p = {a[i].x.to_mont(), a[i].y.to_mont()};
  Affine<F> q = {b[i].x.to_mont(), b[i].y.to_mont()};
  // exercise both the mixed and the full addition: (p as XYZZ) + q, then + infinity via add()
  XYZZ<F> acc = XYZZ<F>::from_affine(p);
  acc.madd(q);
  XYZZ<F> z = XYZZ<F>::inf();
  z.add(acc);
  Affine<F> r = z.to_affine();
  out[i] = {r.x.from_mont(), r.y.from_mont()};
}

// The same sum through the quad-lane operations of ec_quad.cuh (four lanes per pair): a + b by Quad::add,
// and 2*(a + b) - (a + b) folded in through Quad::dbl and a second Quad::add so the doubling path and the
// P + (-P) path of the combination are exercised too:  r = (2s) + (-s) with s = a + b.
template <class F>
__global__ void dbg_point_add_quad_kernel(const Affine<F>* a, const Affine<F>* b, uint64_t n, Affine<F>* out) {
  const uint64_t gt = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t i = gt >> 2;
  const int ql = threadIdx.x & 3;
  const bool live = i < n;
  if (!live) i = n - 1;  // whole warps take part in the shuffles
  Affine<F> p = {a[i].x.to_mont(), a[i].y.to_mont()};
  Affine<F> q = {b[i].x.to_mont(), b[i].y.to_mont()};
  XYZZ<F> s = Quad<F>::add(XYZZ<F>::from_affine(p), XYZZ<F>::from_affine(q), ql);
  XYZZ<F> r = Quad<F>::add(Quad<F>::dbl(s, ql), s.neg(), ql);
  if (live && ql == 0) {
    Affine<F> o = r.to_affine();
    out[i] = {o.x.from_mont(), o.y.from_mont()};
  }
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
    if (prop.major != 10)
      throw std::runtime_error(std::string("libzkp_b200 is built for sm_100a only; device is ") + prop.name);
    CUDA_CHECK(cudaSetDevice(device));
    g_ctx.device = device;
    g_ctx.sm_count = prop.multiProcessorCount;
    CUDA_CHECK(cudaStreamCreateWithFlags(&g_ctx.stream, cudaStreamNonBlocking));
    CUDA_CHECK(cudaStreamCreateWithFlags(&g_ctx.stream2, cudaStreamNonBlocking));
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
    if (offset + n > r->n) throw InvalidArgument("zkp_scalars_download: range out of bounds");
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

int zkp_latency_probe(int mode, double* ns_per_op) {
  return guarded([&](Context& c) {
    if (mode < 0 || mode > 9 || !ns_per_op) throw InvalidArgument("zkp_latency_probe: bad mode");
    DevBuf out;
    out.reserve(64);
    const int iters = 2000;
    latency_probe_kernel<<<1, 32, 0, c.stream>>>(mode, iters / 10, out.as<uint32_t>());
    CUDA_CHECK_LAUNCH();
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
      CUDA_CHECK(cudaEventRecord(g_ev0, c.stream));
      latency_probe_kernel<<<1, 32, 0, c.stream>>>(mode, iters, out.as<uint32_t>());
      CUDA_CHECK_LAUNCH();
      CUDA_CHECK(cudaEventRecord(g_ev1, c.stream));
      CUDA_CHECK(cudaEventSynchronize(g_ev1));
      float ms;
      CUDA_CHECK(cudaEventElapsedTime(&ms, g_ev0, g_ev1));
      if (ms < best) best = ms;
    }
    c.launches += 4;
    *ns_per_op = (double)best * 1e6 / iters;
    out.release();
  });
}

int zkp_imad_peak(int variant, double* gmacs_per_s, double* sm_clock_mhz_effective) {
  return guarded([&](Context& c) {
    if (variant < 0 || variant > 3 || !gmacs_per_s) throw InvalidArgument("zkp_imad_peak: bad variant");
    const int blocks = c.sm_count * 8, threads = 256;
    const int iters = variant == 3 ? 2000 : 4000;
    DevBuf out;
    out.reserve((size_t)blocks * threads * 4);
    auto launch = [&](int it) {
      switch (variant) {
        case 0: imad_peak_kernel<0><<<blocks, threads, 0, c.stream>>>(out.as<uint32_t>(), it, 0x12345677u); break;
        case 1: imad_peak_kernel<1><<<blocks, threads, 0, c.stream>>>(out.as<uint32_t>(), it, 0x12345677u); break;
        case 2: imad_peak_kernel<2><<<blocks, threads, 0, c.stream>>>(out.as<uint32_t>(), it, 0x12345677u); break;
        default: imad_peak_kernel<3><<<blocks, threads, 0, c.stream>>>(out.as<uint32_t>(), it, 0x12345677u); break;
      }
      CUDA_CHECK_LAUNCH();
      c.launches++;
    };
    launch(iters / 10);  // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
      CUDA_CHECK(cudaEventRecord(g_ev0, c.stream));
      launch(iters);
      CUDA_CHECK(cudaEventRecord(g_ev1, c.stream));
      CUDA_CHECK(cudaEventSynchronize(g_ev1));
      float ms;
      CUDA_CHECK(cudaEventElapsedTime(&ms, g_ev0, g_ev1));
      if (ms < best) best = ms;
    }
    double per_thread = variant == 3 ? (double)iters * 2 * 136 : (double)iters * PEAK_UNROLL * 8;
    double macs = per_thread * blocks * threads;
    *gmacs_per_s = macs / (best * 1e-3) / 1e9;
    if (sm_clock_mhz_effective) *sm_clock_mhz_effective = 0.0;  // clocks are sampled by bench.py via nvidia-smi
    out.release();
  });
}

int zkp_dbg_field_op(int field, int op, const uint8_t* a, const uint8_t* b, uint64_t n, uint8_t* out) {
  return guarded([&](Context& c) {
    if (!a || !out || op < 0 || op > 5) throw InvalidArgument("zkp_dbg_field_op: bad argument");
    if (n == 0) return;
    DevBuf da, db, dout;
    da.reserve(n * 32);
    dout.reserve(n * 32);
    CUDA_CHECK(cudaMemcpyAsync(da.p, a, n * 32, cudaMemcpyHostToDevice, c.stream));
    if (b) {
      db.reserve(n * 32);
      CUDA_CHECK(cudaMemcpyAsync(db.p, b, n * 32, cudaMemcpyHostToDevice, c.stream));
    }
    if (field == 0)
      dbg_field_op_kernel<Fp><<<ceil_div(n, 128), 128, 0, c.stream>>>(op, da.as<Fp>(), b ? db.as<Fp>() : nullptr, n,
                                                                     dout.as<Fp>());
    else
      dbg_field_op_kernel<Fr><<<ceil_div(n, 128), 128, 0, c.stream>>>(op, da.as<Fr>(), b ? db.as<Fr>() : nullptr, n,
                                                                     dout.as<Fr>());
    CUDA_CHECK_LAUNCH();
    c.launches++;
    CUDA_CHECK(cudaMemcpyAsync(out, dout.p, n * 32, cudaMemcpyDeviceToHost, c.stream));
    CUDA_CHECK(cudaStreamSynchronize(c.stream));
    da.release();
    db.release();
    dout.release();
  });
}

int zkp_dbg_point_add(int group, const uint8_t* a, const uint8_t* b, uint64_t n, uint8_t* out) {
  return guarded([&](Context& c) {
    if (!a || !b || !out) throw InvalidArgument("zkp_dbg_point_add: null argument");
    if (n == 0) return;
    if (group < 0 || group > 3) throw InvalidArgument("zkp_dbg_point_add: group must be 0 (G1), 1 (G2), 2 / 3 (the same on quads of lanes)");
    size_t sz = (group & 1) == 0 ? 64 : 128;
    DevBuf da, db, dout;
    da.reserve(n * sz);
    db.reserve(n * sz);
    dout.reserve(n * sz);
    CUDA_CHECK(cudaMemcpyAsync(da.p, a, n * sz, cudaMemcpyHostToDevice, c.stream));
    CUDA_CHECK(cudaMemcpyAsync(db.p, b, n * sz, cudaMemcpyHostToDevice, c.stream));
    if (group == 2)
      dbg_point_add_quad_kernel<Fp><<<ceil_div(4 * n, 64), 64, 0, c.stream>>>(da.as<G1Affine>(), db.as<G1Affine>(), n,
                                                                             dout.as<G1Affine>());
    else if (group == 3)
      dbg_point_add_quad_kernel<Fp2><<<ceil_div(4 * n, 64), 64, 0, c.stream>>>(da.as<G2Affine>(), db.as<G2Affine>(), n,
                                                                              dout.as<G2Affine>());
    else if (group == 0)
      dbg_point_add_kernel<Fp><<<ceil_div(n, 64), 64, 0, c.stream>>>(da.as<G1Affine>(), db.as<G1Affine>(), n,
                                                                    dout.as<G1Affine>());
    else
      dbg_point_add_kernel<Fp2><<<ceil_div(n, 64), 64, 0, c.stream>>>(da.as<G2Affine>(), db.as<G2Affine>(), n,
                                                                     dout.as<G2Affine>());
    CUDA_CHECK_LAUNCH();
    c.launches++;
    CUDA_CHECK(cudaMemcpyAsync(out, dout.p, n * sz, cudaMemcpyDeviceToHost, c.stream));
    CUDA_CHECK(cudaStreamSynchronize(c.stream));
    da.release();
    db.release();
    dout.release();
  });
}

}  // extern "C"
