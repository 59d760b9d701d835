// Fr vector / polynomial layer of libzkp_b200 (include/zkp_b200.h): pointwise ops, batch inversion,
// Horner evaluation, NTT-based products, division with remainder, the Groth16 quotient.
//
// Replaces the coefficient-form Python loops of the reference:
//   Polynomial.__add__/__sub__/__mul__/evaluate   /root/reference/zkp/plonk/polynomial.py:85-159
//   poly_div (long division)                      /root/reference/zkp/plonk/polynomial.py:385-435
//   _multiply_polys/_subtract_polys/_div_polys/hxr /root/reference/zkp/groth16/poly_utils.py:17-45,116-125
//   _multiply_vec_matrix                          /root/reference/zkp/groth16/poly_utils.py:52-59
//   compute_accumulator's n-1 divisions           /root/reference/zkp/plonk/permutation.py:120-135
// The reference's schoolbook product is O(n^2) and its long division O(n^2)..O(n^3); here a product
// is three NTTs and a division is a Newton power-series inversion of the reversed divisor
// (quotient) followed by one cyclic product (remainder).  The quotient and remainder of a division
// are unique, so the results are bit-identical with the reference's.
#include <cstring>
#include <vector>
#include "ntt.cuh"
#include "registry.cuh"

namespace zkp {

// ------------------------------------------------------------------ arena of reusable device chunks
struct Arena {
  struct Chunk {
    DevBuf buf;
    bool used = false;
  };
  std::vector<std::unique_ptr<Chunk>> chunks;
  Fr* alloc(uint64_t n_elems) {
    size_t bytes = (size_t)(n_elems ? n_elems : 1) * sizeof(Fr);
    Chunk* best = nullptr;
    for (auto& ch : chunks)
      if (!ch->used && ch->buf.cap >= bytes && (!best || ch->buf.cap < best->buf.cap)) best = ch.get();
    if (!best) {
      for (auto& ch : chunks)
        if (!ch->used && (!best || ch->buf.cap > best->buf.cap)) best = ch.get();
      if (!best) {
        chunks.emplace_back(new Chunk());
        best = chunks.back().get();
      }
      best->buf.reserve(bytes);
    }
    best->used = true;
    return best->buf.as<Fr>();
  }
  // stream-ordered reuse: a freed chunk may be handed out again to work enqueued later on the same stream
  void free(Fr* p) {
    for (auto& ch : chunks)
      if (ch->buf.p == p) ch->used = false;
  }
  void reset() {
    for (auto& ch : chunks) ch->used = false;
  }
};
static Arena g_arena;
struct ArenaScope {
  ~ArenaScope() { g_arena.reset(); }
};

static FrBytes omega_for(uint32_t log_n) {
  FrBytes w;
  memcpy(w.b, FrParams::OMEGA[log_n], 32);
  return w;
}
static uint32_t log2_ceil(uint64_t n) {
  uint32_t l = 0;
  while ((uint64_t(1) << l) < n) l++;
  return l;
}

// ------------------------------------------------------------------ elementwise kernels
#define GRID_1D(n) ceil_div((n), 256), 256, 0, c.stream
#define IDX64 ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x)

// canonical in, canonical out.  op: 0 add, 1 sub, 2 mul, 3 scale by b[0]
__global__ void fr_vec_op_kernel(int op, const Fr* __restrict__ a, const Fr* __restrict__ b, uint64_t n,
                                 Fr* __restrict__ out) {
  uint64_t i = IDX64;
  if (i >= n) return;
  Fr x = a[i], y = b[op == 3 ? 0 : i], r;
  if (op == 0) r = x + y;
  else if (op == 1) r = x - y;
  else r = (x * y) * Fr::r2();  // (xy R^-1)(R^2) R^-1 = xy
  out[i] = r;
}
__global__ void fr_to_mont_kernel(const Fr* __restrict__ in, uint64_t n, Fr* __restrict__ out) {
  uint64_t i = IDX64;
  if (i < n) out[i] = in[i].to_mont();
}
__global__ void fr_from_mont_kernel(const Fr* __restrict__ in, uint64_t n, Fr* __restrict__ out) {
  uint64_t i = IDX64;
  if (i < n) out[i] = in[i].from_mont();
}
__global__ void fr_mul_inplace_kernel(Fr* __restrict__ a, const Fr* __restrict__ b, uint64_t n) {
  uint64_t i = IDX64;
  if (i < n) a[i] = a[i] * b[i];
}
__global__ void fr_sub_inplace_kernel(Fr* __restrict__ a, const Fr* __restrict__ b, uint64_t n) {
  uint64_t i = IDX64;
  if (i < n) a[i] = a[i] - b[i];
}
// out[i] = in[len-1-i] for i < m (zero where the index runs off the front)
__global__ void fr_reverse_kernel(const Fr* __restrict__ in, uint64_t len, uint64_t m, Fr* __restrict__ out) {
  uint64_t i = IDX64;
  if (i < m) out[i] = i < len ? in[len - 1 - i] : Fr::zero();
}
// out[i] = sum_k in[i + k*N], i < N   (reduction modulo x^N - 1)
__global__ void fr_fold_kernel(const Fr* __restrict__ in, uint64_t len, uint64_t N, Fr* __restrict__ out) {
  uint64_t i = IDX64;
  if (i >= N) return;
  Fr acc = Fr::zero();
  for (uint64_t k = i; k < len; k += N) acc = acc + in[k];
  out[i] = acc;
}
// t[i] = (i == 0 ? 2 : 0) - t[i]   (Montgomery)
__global__ void fr_two_minus_kernel(Fr* __restrict__ t, uint64_t n) {
  uint64_t i = IDX64;
  if (i >= n) return;
  Fr two = Fr::one() + Fr::one();
  t[i] = (i == 0 ? two : Fr::zero()) - t[i];
}
__global__ void fr_inv_single_kernel(const Fr* __restrict__ in, Fr* __restrict__ out) {
  if (IDX64 == 0) *out = in->inv();
}
// out[i] = a[i] * (*s)
__global__ void fr_scale_by_dev_kernel(const Fr* __restrict__ a, const Fr* __restrict__ s, uint64_t n,
                                       Fr* __restrict__ out) {
  uint64_t i = IDX64;
  if (i < n) out[i] = a[i] * (*s);
}

// Batch inversion, Montgomery's trick on strided chunks: thread t owns elements t, t+T, t+2T, ...
// (coalesced), one Fermat inversion per BATCH_INV_CHUNK elements.  Zeros map to zero (py_ecc inv(0)=0).
// Input and output in the same form F: canonical (mont == 0) or Montgomery (mont == 1).
static constexpr int BATCH_INV_CHUNK = 32;
__global__ void __launch_bounds__(128) fr_batch_inverse_kernel(const Fr* __restrict__ in, uint64_t n, uint64_t T, int mont,
                                                                Fr* __restrict__ out) {
  uint64_t t = IDX64;
  if (t >= T) return;
  Fr pre[BATCH_INV_CHUNK];
  Fr acc = Fr::one();
  int cnt = 0;
  for (uint64_t i = t; i < n && cnt < BATCH_INV_CHUNK; i += T, cnt++) {
    Fr x = in[i];
    if (!mont) x = x.to_mont();
    pre[cnt] = acc;
    if (!x.is_zero()) acc = acc * x;
  }
  Fr inv = acc.inv();
  for (int k = cnt - 1; k >= 0; k--) {
    uint64_t i = t + (uint64_t)k * T;
    Fr x = in[i];
    if (!mont) x = x.to_mont();
    Fr r = Fr::zero();
    if (!x.is_zero()) {
      r = inv * pre[k];
      inv = inv * x;
    }
    out[i] = mont ? r : r.from_mont();
  }
}

// Horner in two levels: thread j evaluates its chunk of HORNER_CHUNK coefficients at x, then a
// single block combines the partials with x^(HORNER_CHUNK*j).  coeffs canonical, x Montgomery.
static constexpr int HORNER_CHUNK = 64;
__global__ void horner_partial_kernel(const Fr* __restrict__ coeffs, uint64_t n, Fr x, Fr* __restrict__ partial) {
  uint64_t j = IDX64;
  uint64_t beg = j * HORNER_CHUNK;
  if (beg >= n) return;
  uint64_t end = beg + HORNER_CHUNK < n ? beg + HORNER_CHUNK : n;
  Fr acc = Fr::zero();
  for (uint64_t i = end; i > beg; i--) acc = acc * x + coeffs[i - 1];  // canonical coeffs: acc stays canonical
  partial[j] = acc;
}
// single block of 256 threads: out = sum_j partial[j] * xc^j, xc = x^HORNER_CHUNK given as pow2 table
__global__ void horner_combine_kernel(const Fr* __restrict__ partial, uint64_t np, const Fr* __restrict__ xc_pow2,
                                      Fr* __restrict__ out) {
  __shared__ Fr sm[256];
  Fr acc = Fr::zero();
  for (uint64_t j = threadIdx.x; j < np; j += 256) {
    Fr p = partial[j];
    // xc^j from the table
    Fr w = Fr::one();
    bool first = true;
    uint64_t e = j;
    for (int k = 0; e; k++, e >>= 1)
      if (e & 1) {
        if (first) { w = xc_pow2[k]; first = false; }
        else w = w * xc_pow2[k];
      }
    acc = acc + (first ? p : p * w);  // p canonical, w Montgomery -> canonical
  }
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sm[threadIdx.x] = sm[threadIdx.x] + sm[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = sm[0];
}
// tab[k] = x^(HORNER_CHUNK * 2^k), k < 40
__global__ void horner_pow_table_kernel(Fr x, Fr* __restrict__ tab) {
  if (IDX64 != 0) return;
  Fr xc = x;
  for (int k = 1; k < HORNER_CHUNK; k <<= 1) xc = xc.sqr();
  for (int k = 0; k < 40; k++) {
    tab[k] = xc;
    xc = xc.sqr();
  }
}

// out[j] = sum_i vec[i] * mat[i*cols + j]   (canonical in/out)
__global__ void fr_vec_matrix_kernel(const Fr* __restrict__ vec, const Fr* __restrict__ mat, uint64_t rows, uint64_t cols,
                                     Fr* __restrict__ out) {
  uint64_t j = IDX64;
  if (j >= cols) return;
  Fr acc = Fr::zero();
  for (uint64_t i = 0; i < rows; i++) acc = acc + vec[i].to_mont() * mat[i * cols + j];  // (vR)(m)R^-1 = vm
  out[j] = acc;
}

// Division by x^n - 1: q_i = sum_{k>=1} a[i + k n], r_i = a_i + q_i (i < n).  One thread per residue.
__global__ void fr_div_vanishing_kernel(const Fr* __restrict__ a, uint64_t la, uint64_t n, Fr* __restrict__ q,
                                        uint64_t lq, Fr* __restrict__ r) {
  uint64_t i = IDX64;
  if (i >= n) return;
  // walk the residue class from the top down
  uint64_t top = i + ((la - 1 - i) / n) * n;  // largest index == i mod n below la (la > i guaranteed by caller)
  Fr run = Fr::zero();
  for (uint64_t k = top; k >= n + i; k -= n) {
    run = run + a[k];
    q[k - n] = run;
  }
  r[i] = a[i] + run;
  (void)lq;
}


// ------------------------------------------------------------------ exclusive prefix product
// out[0] = 1, out[i] = in[0] * ... * in[i-1]  (canonical in and out).  Tiles of 1024 elements:
// tile products -> single-block scan of the tile products -> per-tile scan.  Replaces running
// products such as tau^i (srs.py:78-82) and the accumulator z (permutation.py:120-135).
static constexpr int PP_THREADS = 256, PP_ITEMS = 4, PP_TILE = PP_THREADS * PP_ITEMS;

// inclusive block scan (Hillis-Steele in shared memory) of one Montgomery value per thread
__device__ __forceinline__ Fr block_inclusive_product(Fr v, Fr* sm) {
  sm[threadIdx.x] = v;
  __syncthreads();
  for (int o = 1; o < PP_THREADS; o <<= 1) {
    Fr t = v;
    if ((int)threadIdx.x >= o) t = sm[threadIdx.x - o] * v;
    __syncthreads();
    v = t;
    sm[threadIdx.x] = v;
    __syncthreads();
  }
  return v;
}

__global__ void __launch_bounds__(PP_THREADS) pp_tile_products_kernel(const Fr* __restrict__ in, uint64_t n,
                                                                       Fr* __restrict__ tile_prod) {
  __shared__ Fr sm[PP_THREADS];
  uint64_t base = (uint64_t)blockIdx.x * PP_TILE + (uint64_t)threadIdx.x * PP_ITEMS;
  Fr p = Fr::one();
#pragma unroll
  for (int k = 0; k < PP_ITEMS; k++)
    if (base + k < n) p = p * in[base + k].to_mont();
  Fr inc = block_inclusive_product(p, sm);
  if (threadIdx.x == PP_THREADS - 1) tile_prod[blockIdx.x] = inc;
}

// single block: tile_prod[i] <- exclusive prefix product (Montgomery)
__global__ void __launch_bounds__(PP_THREADS) pp_tile_scan_kernel(Fr* __restrict__ tile_prod, uint64_t ntiles) {
  __shared__ Fr sm[PP_THREADS];
  __shared__ Fr carry;
  if (threadIdx.x == 0) carry = Fr::one();
  __syncthreads();
  for (uint64_t base = 0; base < ntiles; base += PP_THREADS) {
    uint64_t i = base + threadIdx.x;
    Fr v = i < ntiles ? tile_prod[i] : Fr::one();
    Fr inc = block_inclusive_product(v, sm);
    Fr c = carry;
    Fr prev = threadIdx.x ? sm[threadIdx.x - 1] : Fr::one();
    __syncthreads();
    if (i < ntiles) tile_prod[i] = c * prev;
    if (threadIdx.x == PP_THREADS - 1) carry = c * inc;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(PP_THREADS) pp_apply_kernel(const Fr* __restrict__ in, uint64_t n,
                                                               const Fr* __restrict__ tile_off, Fr* __restrict__ out) {
  __shared__ Fr sm[PP_THREADS];
  uint64_t base = (uint64_t)blockIdx.x * PP_TILE + (uint64_t)threadIdx.x * PP_ITEMS;
  Fr x[PP_ITEMS];
  Fr p = Fr::one();
#pragma unroll
  for (int k = 0; k < PP_ITEMS; k++) {
    x[k] = (base + k < n) ? in[base + k].to_mont() : Fr::one();
    p = p * x[k];
  }
  block_inclusive_product(p, sm);
  Fr run = tile_off[blockIdx.x];
  if (threadIdx.x) run = run * sm[threadIdx.x - 1];
#pragma unroll
  for (int k = 0; k < PP_ITEMS; k++) {
    if (base + k < n) out[base + k] = run.from_mont();
    run = run * x[k];
  }
}

// ------------------------------------------------------------------ device-level polynomial ops (Montgomery)
// out[0..out_len) = (a * b)[0..out_len); a, b, out device Montgomery; out must not alias a or b.
static int poly_mul_dev(Context& c, const Fr* a, uint64_t la, const Fr* b, uint64_t lb, Fr* out, uint64_t out_len) {
  int launches = 0;
  uint64_t full = la + lb - 1;
  uint32_t lg = log2_ceil(full);
  uint64_t N = uint64_t(1) << lg;
  Fr* fa = g_arena.alloc(N);
  Fr* fb = (a == b && la == lb) ? fa : g_arena.alloc(N);
  Fr* scratch = g_arena.alloc(N);
  CUDA_CHECK(cudaMemcpyAsync(fa, a, la * sizeof(Fr), cudaMemcpyDeviceToDevice, c.stream));
  CUDA_CHECK(cudaMemsetAsync(fa + la, 0, (N - la) * sizeof(Fr), c.stream));
  FrBytes w = omega_for(lg);
  launches += ntt_device(c, fa, scratch, lg, w, false, nullptr);
  if (fb != fa) {
    CUDA_CHECK(cudaMemcpyAsync(fb, b, lb * sizeof(Fr), cudaMemcpyDeviceToDevice, c.stream));
    CUDA_CHECK(cudaMemsetAsync(fb + lb, 0, (N - lb) * sizeof(Fr), c.stream));
    launches += ntt_device(c, fb, scratch, lg, w, false, nullptr);
  }
  fr_mul_inplace_kernel<<<GRID_1D(N)>>>(fa, fb, N);
  CUDA_CHECK_LAUNCH();
  launches++;
  launches += ntt_device(c, fa, scratch, lg, w, true, nullptr);
  uint64_t take = out_len < full ? out_len : full;
  CUDA_CHECK(cudaMemcpyAsync(out, fa, take * sizeof(Fr), cudaMemcpyDeviceToDevice, c.stream));
  if (out_len > full) CUDA_CHECK(cudaMemsetAsync(out + full, 0, (out_len - full) * sizeof(Fr), c.stream));
  if (fb != fa) g_arena.free(fb);
  g_arena.free(fa);
  g_arena.free(scratch);
  return launches;
}

// g = f^-1 mod x^m by Newton iteration (f[0] != 0); f has lf coefficients.  g: m elements.
static int series_inverse_dev(Context& c, const Fr* f, uint64_t lf, uint64_t m, Fr* g) {
  int launches = 0;
  fr_inv_single_kernel<<<1, 32, 0, c.stream>>>(f, g);
  CUDA_CHECK_LAUNCH();
  launches++;
  uint64_t k = 1;
  Fr* t = g_arena.alloc(2 * m + 2);
  Fr* g2 = g_arena.alloc(2 * m + 2);
  while (k < m) {
    uint64_t k2 = 2 * k < m ? 2 * k : m;
    uint64_t fl = lf < k2 ? lf : k2;
    launches += poly_mul_dev(c, f, fl, g, k, t, k2);  // t = f*g mod x^k2
    fr_two_minus_kernel<<<GRID_1D(k2)>>>(t, k2);
    CUDA_CHECK_LAUNCH();
    launches++;
    launches += poly_mul_dev(c, g, k, t, k2, g2, k2);  // g2 = g*(2 - f g) mod x^k2
    CUDA_CHECK(cudaMemcpyAsync(g, g2, k2 * sizeof(Fr), cudaMemcpyDeviceToDevice, c.stream));
    k = k2;
  }
  return launches;
}

// a = b*q + r.  a (la), b (lb >= 1, b[lb-1] != 0), la >= lb.  q: la-lb+1 elements, r: lb-1 elements.
// All device Montgomery.  vanishing: b is x^(lb-1) - 1 (fast path).
// inv_cache: optional persistent buffer holding rev(b)^-1 mod x^m from an earlier call with the same
// divisor (the Groth16 Z(x) is fixed per circuit); *inv_cached tells whether it is already filled.
static int poly_divmod_dev(Context& c, const Fr* a, uint64_t la, const Fr* b, uint64_t lb, Fr* q, Fr* r, bool vanishing,
                           Fr* inv_cache = nullptr, bool* inv_cached = nullptr) {
  int launches = 0;
  uint64_t m = la - lb + 1;
  if (vanishing && lb >= 2) {
    uint64_t n = lb - 1;
    fr_div_vanishing_kernel<<<GRID_1D(n)>>>(a, la, n, q, m, r);
    CUDA_CHECK_LAUNCH();
    return 1;
  }
  Fr* ra = g_arena.alloc(m);
  uint64_t lrb = lb < m ? lb : m;
  Fr* rb = g_arena.alloc(lrb);
  Fr* g = inv_cache ? inv_cache : g_arena.alloc(m);
  Fr* qr = g_arena.alloc(m);
  fr_reverse_kernel<<<GRID_1D(m)>>>(a, la, m, ra);
  CUDA_CHECK_LAUNCH();
  fr_reverse_kernel<<<GRID_1D(lrb)>>>(b, lb, lrb, rb);
  CUDA_CHECK_LAUNCH();
  launches += 2;
  if (!(inv_cached && *inv_cached)) {
    launches += series_inverse_dev(c, rb, lrb, m, g);
    if (inv_cached) *inv_cached = true;
  }
  launches += poly_mul_dev(c, ra, m, g, m, qr, m);
  fr_reverse_kernel<<<GRID_1D(m)>>>(qr, m, m, q);
  CUDA_CHECK_LAUNCH();
  launches++;
  if (lb >= 2) {
    // r = (a - b q) mod (x^N - 1), N >= lb-1: exact because deg r < lb-1 <= N
    uint64_t lr = lb - 1;
    uint32_t lg = log2_ceil(lr);
    uint64_t N = uint64_t(1) << lg;
    Fr* fa = g_arena.alloc(N);
    Fr* fb = g_arena.alloc(N);
    Fr* fq = g_arena.alloc(N);
    Fr* scratch = g_arena.alloc(N);
    fr_fold_kernel<<<GRID_1D(N)>>>(a, la, N, fa);
    CUDA_CHECK_LAUNCH();
    fr_fold_kernel<<<GRID_1D(N)>>>(b, lb, N, fb);
    CUDA_CHECK_LAUNCH();
    fr_fold_kernel<<<GRID_1D(N)>>>(q, m, N, fq);
    CUDA_CHECK_LAUNCH();
    launches += 3;
    FrBytes w = omega_for(lg);
    launches += ntt_device(c, fb, scratch, lg, w, false, nullptr);
    launches += ntt_device(c, fq, scratch, lg, w, false, nullptr);
    fr_mul_inplace_kernel<<<GRID_1D(N)>>>(fb, fq, N);
    CUDA_CHECK_LAUNCH();
    launches += ntt_device(c, fb, scratch, lg, w, true, nullptr);
    fr_sub_inplace_kernel<<<GRID_1D(N)>>>(fa, fb, N);
    CUDA_CHECK_LAUNCH();
    launches += 2;
    CUDA_CHECK(cudaMemcpyAsync(r, fa, lr * sizeof(Fr), cudaMemcpyDeviceToDevice, c.stream));
  }
  return launches;
}

// upload canonical host bytes -> device Montgomery
static Fr* upload_mont(Context& c, const uint8_t* host, uint64_t n, int* launches) {
  Fr* d = g_arena.alloc(n);
  if (n) {
    CUDA_CHECK(cudaMemcpyAsync(d, host, n * 32, cudaMemcpyHostToDevice, c.stream));
    fr_to_mont_kernel<<<GRID_1D(n)>>>(d, n, d);
    CUDA_CHECK_LAUNCH();
    (*launches)++;
  }
  return d;
}
static void download_canon(Context& c, Fr* d, uint64_t n, uint8_t* host, int* launches) {
  if (!n) return;
  fr_from_mont_kernel<<<GRID_1D(n)>>>(d, n, d);
  CUDA_CHECK_LAUNCH();
  (*launches)++;
  CUDA_CHECK(cudaMemcpyAsync(host, d, n * 32, cudaMemcpyDeviceToHost, c.stream));
}

static bool host_is_vanishing(const uint8_t* b, uint64_t lb) {
  if (lb < 2) return false;
  uint8_t minus_one[32], one[32] = {1};
  memcpy(minus_one, FrParams::MOD, 32);
  minus_one[0] -= 1;  // r - 1: low byte of r is 0x01
  if (memcmp(b, minus_one, 32) != 0 || memcmp(b + (lb - 1) * 32, one, 32) != 0) return false;
  for (uint64_t i = 1; i + 1 < lb; i++) {
    const uint64_t* p = reinterpret_cast<const uint64_t*>(b + i * 32);
    if (p[0] | p[1] | p[2] | p[3]) return false;
  }
  return true;
}
static bool host_is_zero32(const uint8_t* p) {
  for (int i = 0; i < 32; i++)
    if (p[i]) return false;
  return true;
}

}  // namespace zkp

using namespace zkp;

extern "C" {

int zkp_fr_vec_op(int op, const uint8_t* a, const uint8_t* b, uint64_t n, uint8_t* out) {
  return guarded([&](Context& c) {
    if (op < 0 || op > 3 || (n && (!a || !b || !out))) throw InvalidArgument("zkp_fr_vec_op: bad argument");
    if (!n) return;
    ArenaScope scope;
    Fr* da = g_arena.alloc(n);
    uint64_t nb = op == 3 ? 1 : n;
    Fr* db = g_arena.alloc(nb);
    CUDA_CHECK(cudaMemcpyAsync(da, a, n * 32, cudaMemcpyHostToDevice, c.stream));
    CUDA_CHECK(cudaMemcpyAsync(db, b, nb * 32, cudaMemcpyHostToDevice, c.stream));
    fr_vec_op_kernel<<<GRID_1D(n)>>>(op, da, db, n, da);
    CUDA_CHECK_LAUNCH();
    c.launches++;
    CUDA_CHECK(cudaMemcpyAsync(out, da, n * 32, cudaMemcpyDeviceToHost, c.stream));
    CUDA_CHECK(cudaStreamSynchronize(c.stream));
  });
}

int zkp_fr_batch_inverse(const uint8_t* a, uint64_t n, uint8_t* out) {
  return guarded([&](Context& c) {
    if (n && (!a || !out)) throw InvalidArgument("zkp_fr_batch_inverse: null argument");
    if (!n) return;
    ArenaScope scope;
    Fr* da = g_arena.alloc(n);
    Fr* dout = g_arena.alloc(n);
    CUDA_CHECK(cudaMemcpyAsync(da, a, n * 32, cudaMemcpyHostToDevice, c.stream));
    uint64_t T = (n + BATCH_INV_CHUNK - 1) / BATCH_INV_CHUNK;
    fr_batch_inverse_kernel<<<ceil_div(T, 128), 128, 0, c.stream>>>(da, n, T, 0, dout);
    CUDA_CHECK_LAUNCH();
    c.launches++;
    CUDA_CHECK(cudaMemcpyAsync(out, dout, n * 32, cudaMemcpyDeviceToHost, c.stream));
    CUDA_CHECK(cudaStreamSynchronize(c.stream));
  });
}

int zkp_fr_prefix_product(const uint8_t* a, uint64_t n, uint8_t* out) {
  return guarded([&](Context& c) {
    if (n && (!a || !out)) throw InvalidArgument("zkp_fr_prefix_product: null argument");
    if (!n) return;
    ArenaScope scope;
    uint64_t ntiles = (n + PP_TILE - 1) / PP_TILE;
    Fr* da = g_arena.alloc(n);
    Fr* dout = g_arena.alloc(n);
    Fr* tiles = g_arena.alloc(ntiles);
    CUDA_CHECK(cudaMemcpyAsync(da, a, n * 32, cudaMemcpyHostToDevice, c.stream));
    pp_tile_products_kernel<<<(unsigned)ntiles, PP_THREADS, 0, c.stream>>>(da, n, tiles);
    CUDA_CHECK_LAUNCH();
    pp_tile_scan_kernel<<<1, PP_THREADS, 0, c.stream>>>(tiles, ntiles);
    CUDA_CHECK_LAUNCH();
    pp_apply_kernel<<<(unsigned)ntiles, PP_THREADS, 0, c.stream>>>(da, n, tiles, dout);
    CUDA_CHECK_LAUNCH();
    c.launches += 3;
    CUDA_CHECK(cudaMemcpyAsync(out, dout, n * 32, cudaMemcpyDeviceToHost, c.stream));
    CUDA_CHECK(cudaStreamSynchronize(c.stream));
  });
}

int zkp_fr_poly_eval(const uint8_t* coeffs, uint64_t n, const uint8_t x[32], uint8_t out[32]) {
  return guarded([&](Context& c) {
    if (!x || !out || (n && !coeffs)) throw InvalidArgument("zkp_fr_poly_eval: null argument");
    if (!n) {
      memset(out, 0, 32);
      return;
    }
    ArenaScope scope;
    Fr* dc = g_arena.alloc(n);
    uint64_t np = (n + HORNER_CHUNK - 1) / HORNER_CHUNK;
    Fr* partial = g_arena.alloc(np);
    Fr* tab = g_arena.alloc(41);
    CUDA_CHECK(cudaMemcpyAsync(dc, coeffs, n * 32, cudaMemcpyHostToDevice, c.stream));
    // x -> Montgomery on the host side of the launch: pass canonical, convert in a tiny kernel
    Fr xc;
    memcpy(xc.v, x, 32);
    Fr* dx = tab + 40;
    CUDA_CHECK(cudaMemcpyAsync(dx, &xc, 32, cudaMemcpyHostToDevice, c.stream));
    fr_to_mont_kernel<<<1, 32, 0, c.stream>>>(dx, 1, dx);
    CUDA_CHECK_LAUNCH();
    Fr xm;
    CUDA_CHECK(cudaMemcpyAsync(&xm, dx, 32, cudaMemcpyDeviceToHost, c.stream));
    CUDA_CHECK(cudaStreamSynchronize(c.stream));
    horner_pow_table_kernel<<<1, 32, 0, c.stream>>>(xm, tab);
    CUDA_CHECK_LAUNCH();
    horner_partial_kernel<<<GRID_1D(np)>>>(dc, n, xm, partial);
    CUDA_CHECK_LAUNCH();
    horner_combine_kernel<<<1, 256, 0, c.stream>>>(partial, np, tab, dx);
    CUDA_CHECK_LAUNCH();
    c.launches += 4;
    CUDA_CHECK(cudaMemcpyAsync(out, dx, 32, cudaMemcpyDeviceToHost, c.stream));
    CUDA_CHECK(cudaStreamSynchronize(c.stream));
  });
}

int zkp_fr_vec_matrix(const uint8_t* vec, const uint8_t* mat, uint64_t rows, uint64_t cols, uint8_t* out) {
  return guarded([&](Context& c) {
    if (!rows || !cols) return;
    if (!vec || !mat || !out) throw InvalidArgument("zkp_fr_vec_matrix: null argument");
    ArenaScope scope;
    Fr* dv = g_arena.alloc(rows);
    Fr* dm = g_arena.alloc(rows * cols);
    Fr* dout = g_arena.alloc(cols);
    CUDA_CHECK(cudaMemcpyAsync(dv, vec, rows * 32, cudaMemcpyHostToDevice, c.stream));
    CUDA_CHECK(cudaMemcpyAsync(dm, mat, rows * cols * 32, cudaMemcpyHostToDevice, c.stream));
    fr_vec_matrix_kernel<<<GRID_1D(cols)>>>(dv, dm, rows, cols, dout);
    CUDA_CHECK_LAUNCH();
    c.launches++;
    CUDA_CHECK(cudaMemcpyAsync(out, dout, cols * 32, cudaMemcpyDeviceToHost, c.stream));
    CUDA_CHECK(cudaStreamSynchronize(c.stream));
  });
}

int zkp_fr_poly_mul(const uint8_t* a, uint64_t a_len, const uint8_t* b, uint64_t b_len, uint8_t* out) {
  return guarded([&](Context& c) {
    if (!a_len || !b_len || !a || !b || !out) throw InvalidArgument("zkp_fr_poly_mul: empty or null operand");
    ArenaScope scope;
    int launches = 0;
    Fr* da = upload_mont(c, a, a_len, &launches);
    Fr* db = upload_mont(c, b, b_len, &launches);
    uint64_t lo = a_len + b_len - 1;
    Fr* dout = g_arena.alloc(lo);
    launches += poly_mul_dev(c, da, a_len, db, b_len, dout, lo);
    download_canon(c, dout, lo, out, &launches);
    c.launches += launches;
    CUDA_CHECK(cudaStreamSynchronize(c.stream));
  });
}

int zkp_fr_poly_divmod(const uint8_t* a, uint64_t a_len, const uint8_t* b, uint64_t b_len, uint8_t* q_out,
                       uint8_t* r_out) {
  return guarded([&](Context& c) {
    if (!a || !b || !q_out || !b_len || a_len < b_len) throw InvalidArgument("zkp_fr_poly_divmod: need a_len >= b_len >= 1");
    if (b_len > 1 && !r_out) throw InvalidArgument("zkp_fr_poly_divmod: null remainder buffer");
    if (host_is_zero32(b + (b_len - 1) * 32)) throw InvalidArgument("zkp_fr_poly_divmod: leading coefficient of the divisor is zero");
    ArenaScope scope;
    int launches = 0;
    bool van = host_is_vanishing(b, b_len);
    Fr* da = upload_mont(c, a, a_len, &launches);
    Fr* db = upload_mont(c, b, b_len, &launches);
    uint64_t m = a_len - b_len + 1;
    Fr* dq = g_arena.alloc(m);
    Fr* dr = g_arena.alloc(b_len);
    launches += poly_divmod_dev(c, da, a_len, db, b_len, dq, dr, van);
    download_canon(c, dq, m, q_out, &launches);
    download_canon(c, dr, b_len - 1, r_out, &launches);
    c.launches += launches;
    CUDA_CHECK(cudaStreamSynchronize(c.stream));
  });
}

// ------------------------------------------------------------------ device-resident Fr vectors (handles)
struct DivisorCache {
  uint64_t z_handle = 0, z_len = 0, m = 0;
  DevBuf inv;
  bool filled = false;
};
static DivisorCache g_div_cache;

int zkp_scalars_alloc(uint64_t n, uint64_t* handle) {
  return guarded([&](Context& c) {
    if (!handle) throw InvalidArgument("zkp_scalars_alloc: null handle");
    auto r = std::make_unique<Resource>();
    r->kind = HandleKind::Scalars;
    r->n = n;
    r->buf.reserve_pooled(n ? n * 32 : 32);
    CUDA_CHECK(cudaMemsetAsync(r->buf.p, 0, n ? n * 32 : 32, c.stream));
    CUDA_CHECK(cudaStreamSynchronize(c.stream));
    *handle = registry().put(std::move(r));
  });
}

int zkp_scalars_copy(uint64_t dst, uint64_t dst_off, uint64_t src, uint64_t src_off, uint64_t n) {
  return guarded([&](Context& c) {
    Resource* d = need(dst, HandleKind::Scalars, "zkp_scalars_copy");
    Resource* s = need(src, HandleKind::Scalars, "zkp_scalars_copy");
    if (dst_off + n > d->n || src_off + n > s->n) throw InvalidArgument("zkp_scalars_copy: range out of bounds");
    if (n) CUDA_CHECK(cudaMemcpyAsync(d->buf.as<Fr>() + dst_off, s->buf.as<Fr>() + src_off, n * 32,
                                     cudaMemcpyDeviceToDevice, c.stream));
  });
}

int zkp_scalars_upload(uint64_t dst, uint64_t dst_off, const uint8_t* host, uint64_t n) {
  return guarded([&](Context& c) {
    Resource* d = need(dst, HandleKind::Scalars, "zkp_scalars_upload");
    if (dst_off + n > d->n || (n && !host)) throw InvalidArgument("zkp_scalars_upload: bad range or null source");
    if (n) CUDA_CHECK(cudaMemcpyAsync(d->buf.as<Fr>() + dst_off, host, n * 32, cudaMemcpyHostToDevice, c.stream));
    CUDA_CHECK(cudaStreamSynchronize(c.stream));
  });
}

int zkp_scalars_scale(uint64_t h, uint64_t off, uint64_t n, const uint8_t k[32]) {
  return guarded([&](Context& c) {
    Resource* d = need(h, HandleKind::Scalars, "zkp_scalars_scale");
    if (off + n > d->n || !k) throw InvalidArgument("zkp_scalars_scale: bad range or null factor");
    if (!n) return;
    ArenaScope scope;
    Fr* dk = g_arena.alloc(1);
    CUDA_CHECK(cudaMemcpyAsync(dk, k, 32, cudaMemcpyHostToDevice, c.stream));
    fr_vec_op_kernel<<<GRID_1D(n)>>>(3, d->buf.as<Fr>() + off, dk, n, d->buf.as<Fr>() + off);
    CUDA_CHECK_LAUNCH();
    c.launches++;
  });
}

int zkp_fr_poly_eval_dev(uint64_t h, uint64_t off, uint64_t n, const uint8_t x[32], uint8_t out[32]) {
  return guarded([&](Context& c) {
    Resource* d = need(h, HandleKind::Scalars, "zkp_fr_poly_eval_dev");
    if (off + n > d->n || !x || !out) throw InvalidArgument("zkp_fr_poly_eval_dev: bad argument");
    if (!n) {
      memset(out, 0, 32);
      return;
    }
    ArenaScope scope;
    uint64_t np = (n + HORNER_CHUNK - 1) / HORNER_CHUNK;
    Fr* partial = g_arena.alloc(np);
    Fr* tab = g_arena.alloc(41);
    Fr* dx = tab + 40;
    CUDA_CHECK(cudaMemcpyAsync(dx, x, 32, cudaMemcpyHostToDevice, c.stream));
    fr_to_mont_kernel<<<1, 32, 0, c.stream>>>(dx, 1, dx);
    CUDA_CHECK_LAUNCH();
    Fr xm;
    CUDA_CHECK(cudaMemcpyAsync(&xm, dx, 32, cudaMemcpyDeviceToHost, c.stream));
    CUDA_CHECK(cudaStreamSynchronize(c.stream));
    horner_pow_table_kernel<<<1, 32, 0, c.stream>>>(xm, tab);
    CUDA_CHECK_LAUNCH();
    horner_partial_kernel<<<GRID_1D(np)>>>(d->buf.as<Fr>() + off, n, xm, partial);
    CUDA_CHECK_LAUNCH();
    horner_combine_kernel<<<1, 256, 0, c.stream>>>(partial, np, tab, dx);
    CUDA_CHECK_LAUNCH();
    c.launches += 4;
    CUDA_CHECK(cudaMemcpyAsync(out, dx, 32, cudaMemcpyDeviceToHost, c.stream));
    CUDA_CHECK(cudaStreamSynchronize(c.stream));
  });
}

// hxr on device-resident coefficient vectors (the 2^20 configuration needs
// coefficient-level entry points).  a, b, c: `len` canonical coefficients each; z: z_len coefficients of
// the (fixed) divisor.  Creates two new scalar handles: the quotient (2*len - z_len coefficients) and
// the remainder (z_len - 1).  The power-series inverse of the reversed divisor is cached per z handle.
int zkp_groth16_quotient_dev(uint64_t a, uint64_t b, uint64_t cc, uint64_t len, uint64_t z, uint64_t z_len,
                             uint64_t* h_out, uint64_t* rem_out) {
  return guarded([&](Context& c) {
    Resource* ra = need(a, HandleKind::Scalars, "zkp_groth16_quotient_dev");
    Resource* rb = need(b, HandleKind::Scalars, "zkp_groth16_quotient_dev");
    Resource* rc = need(cc, HandleKind::Scalars, "zkp_groth16_quotient_dev");
    Resource* rz = need(z, HandleKind::Scalars, "zkp_groth16_quotient_dev");
    if (!h_out || !rem_out || !len || z_len < 2) throw InvalidArgument("zkp_groth16_quotient_dev: bad argument");
    if (len > ra->n || len > rb->n || len > rc->n || z_len > rz->n) throw InvalidArgument("zkp_groth16_quotient_dev: length exceeds a vector");
    uint64_t lp = 2 * len - 1;
    if (lp < z_len) throw InvalidArgument("zkp_groth16_quotient_dev: divisor longer than the product");
    ArenaScope scope;
    int launches = 0;
    auto to_mont_copy = [&](Resource* r, uint64_t n) {
      Fr* d = g_arena.alloc(n);
      fr_to_mont_kernel<<<GRID_1D(n)>>>(r->buf.as<Fr>(), n, d);
      CUDA_CHECK_LAUNCH();
      launches++;
      return d;
    };
    Fr* da = to_mont_copy(ra, len);
    Fr* db = to_mont_copy(rb, len);
    Fr* dc = to_mont_copy(rc, len);
    Fr* dz = to_mont_copy(rz, z_len);
    Fr* dp = g_arena.alloc(lp);
    launches += poly_mul_dev(c, da, len, db, len, dp, lp);
    fr_sub_inplace_kernel<<<GRID_1D(len)>>>(dp, dc, len);
    CUDA_CHECK_LAUNCH();
    launches++;
    uint64_t m = lp - z_len + 1;
    if (g_div_cache.z_handle != z || g_div_cache.z_len != z_len || g_div_cache.m != m) {
      g_div_cache.z_handle = z;
      g_div_cache.z_len = z_len;
      g_div_cache.m = m;
      g_div_cache.filled = false;
      g_div_cache.inv.reserve(m * sizeof(Fr));
    }
    auto hq = std::make_unique<Resource>();
    hq->kind = HandleKind::Scalars;
    hq->n = m;
    hq->buf.reserve_pooled(m * 32);
    auto hr = std::make_unique<Resource>();
    hr->kind = HandleKind::Scalars;
    hr->n = z_len - 1;
    hr->buf.reserve_pooled(z_len * 32);
    launches += poly_divmod_dev(c, dp, lp, dz, z_len, hq->buf.as<Fr>(), hr->buf.as<Fr>(), false,
                                g_div_cache.inv.as<Fr>(), &g_div_cache.filled);
    fr_from_mont_kernel<<<GRID_1D(m)>>>(hq->buf.as<Fr>(), m, hq->buf.as<Fr>());
    CUDA_CHECK_LAUNCH();
    fr_from_mont_kernel<<<GRID_1D(z_len - 1)>>>(hr->buf.as<Fr>(), z_len - 1, hr->buf.as<Fr>());
    CUDA_CHECK_LAUNCH();
    launches += 2;
    c.launches += launches;
    CUDA_CHECK(cudaStreamSynchronize(c.stream));
    *h_out = registry().put(std::move(hq));
    *rem_out = registry().put(std::move(hr));
  });
}

int zkp_groth16_quotient(const uint8_t* a, const uint8_t* b, const uint8_t* cc, uint64_t len, const uint8_t* z,
                         uint64_t z_len, uint8_t* h_out, uint8_t* rem_out) {
  return guarded([&](Context& c) {
    if (!a || !b || !cc || !z || !h_out || !len || !z_len) throw InvalidArgument("zkp_groth16_quotient: null/empty argument");
    uint64_t lp = 2 * len - 1;
    if (lp < z_len) throw InvalidArgument("zkp_groth16_quotient: divisor longer than the product");
    if (z_len > 1 && !rem_out) throw InvalidArgument("zkp_groth16_quotient: null remainder buffer");
    if (host_is_zero32(z + (z_len - 1) * 32)) throw InvalidArgument("zkp_groth16_quotient: leading coefficient of Z is zero");
    ArenaScope scope;
    int launches = 0;
    Fr* da = upload_mont(c, a, len, &launches);
    Fr* db = upload_mont(c, b, len, &launches);
    Fr* dc = upload_mont(c, cc, len, &launches);
    Fr* dz = upload_mont(c, z, z_len, &launches);
    Fr* dp = g_arena.alloc(lp);
    launches += poly_mul_dev(c, da, len, db, len, dp, lp);
    fr_sub_inplace_kernel<<<GRID_1D(len)>>>(dp, dc, len);  // P = a*b - c
    CUDA_CHECK_LAUNCH();
    launches++;
    uint64_t m = lp - z_len + 1;
    Fr* dq = g_arena.alloc(m);
    Fr* dr = g_arena.alloc(z_len);
    launches += poly_divmod_dev(c, dp, lp, dz, z_len, dq, dr, host_is_vanishing(z, z_len));
    download_canon(c, dq, m, h_out, &launches);
    download_canon(c, dr, z_len - 1, rem_out, &launches);
    c.launches += launches;
    CUDA_CHECK(cudaStreamSynchronize(c.stream));
  });
}

}  // extern "C"
