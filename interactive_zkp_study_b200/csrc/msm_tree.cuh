// Bucket accumulation in AFFINE coordinates with shared inversions (Montgomery's trick), as a pairwise tree.
//
// Replaces, for buckets of up to TREE_MAX entries, the serial XYZZ chain of msm_accumulate_kernel: the inner loops
// of kzg.commit (/root/reference/zkp/plonk/kzg.py:59-67) and proof_a/b/c
// (/root/reference/zkp/groth16/proving.py:27-31,39-43,56-60,66-73) are sums of points, and the affine chord rule
//     lambda = (y2 - y1) / (x2 - x1),  x3 = lambda^2 - x1 - x2,  y3 = lambda (x1 - x3) - y1
// costs 2 products + 1 squaring once 1 / (x2 - x1) is known.  A block shares ONE inversion over thousands of
// additions: per addition 1 product going up (running product of the denominators), 2 coming down (this
// denominator's inverse, the running inverse) -- 5 products + 1 squaring = 788 limb-MACs against 8 + 2 = 1304 for
// the XYZZ mixed addition.  The additions of a shared inversion must be independent of each other, so a bucket is
// summed as a tree: round r adds entries 2j and 2j+1 of every bucket that still holds m_r = ceil(m_0 / 2^r) >= 2
// points and writes entry j of the next round; an odd last entry is copied.  The sum of a bucket is a group
// element, so its affine form does not depend on the order: results stay bit-identical.
//
// Schedule (no per-round scans, no idle lanes):
//   * buckets are counting-sorted by entry count, descending (`order`: bucket | m_0 << 23, first entry).  In that
//     order "bucket still has an addition j in round r" is a prefix: rank < G[(2j+1) 2^r + 1], with G[x] the number
//     of buckets holding >= x entries.  The work of a round is the union over j of those prefixes; item id ->
//     (j, rank) is a search in a table of <= 256 row starts (`plan`, built once per MSM from the count histogram).
//   * round r of bucket b reads its entries at base_r(b) .. and writes at base_{r+1}(b) = (base_r(b) + b) >> 1 in
//     the other buffer (base_0 = the bucket's offset in `sorted`); the regions of consecutive buckets never
//     overlap (floor((x + m + 1) / 2) - floor(x / 2) >= ceil(m / 2)), so no offsets are ever recomputed.
//   * the addition that leaves a bucket with one point writes it to `buckets` (XYZZ with ZZ = ZZZ = 1).
//   * a round costs one inversion LATENCY per block (binary Euclid on one thread, ~60 us), and every round halves
//     the work while that cost stays: after K rounds (default 2: 75 % of the additions) the m_K = ceil(m_0 / 2^K)
//     points left in a bucket are summed as one XYZZ chain by msm_tree_finish_kernel, one thread per bucket in the
//     same descending order (equal chain lengths across a warp).
// Buckets above TREE_MAX entries (skewed scalars, dense shards) keep the XYZZ path: the task kernels only see them.
#pragma once

namespace zkp {

static constexpr uint32_t TREE_MAX = 511;          // m_0 fits 9 bits beside a 23-bit bucket index
static constexpr int TREE_ROUNDS = 9;              // ceil(511 / 2^9) = 1
static constexpr uint32_t TREE_JMAX = 256;         // rows (pair index j) of a round: 2j < m_r <= 511
static constexpr uint32_t TREE_REC = 4 + 2 * (TREE_JMAX + 1);  // per round: NA, NC, light buckets, -, add_start[257], copy_start[257]
static constexpr uint32_t TREE_BUCKET_BITS = 23;

// ---- plan, step 1: histogram of the bucket sizes 1..TREE_MAX; empty buckets become infinity here
template <class F>
__global__ void msm_tree_hist_kernel(const uint32_t* __restrict__ offsets, uint32_t nbuckets, uint32_t* __restrict__ hist,
                                     XYZZ<F>* __restrict__ buckets) {
  __shared__ uint32_t h[TREE_MAX + 1];
  for (uint32_t k = threadIdx.x; k <= TREE_MAX; k += blockDim.x) h[k] = 0;
  __syncthreads();
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < nbuckets) {
    uint32_t cnt = offsets[b + 1] - offsets[b];
    if (cnt == 0) buckets[b] = XYZZ<F>::inf();
    else if (cnt <= TREE_MAX) atomicAdd(&h[cnt], 1u);
  }
  __syncthreads();
  for (uint32_t k = threadIdx.x; k <= TREE_MAX; k += blockDim.x)
    if (h[k]) atomicAdd(&hist[k], h[k]);
}

// ---- plan, step 2 (one block of 256 threads): G[x] = #buckets with x <= m_0 <= TREE_MAX, the emit cursors
// (a bucket of m entries gets a rank in [G[m+1], G[m])), and per round the row starts of additions and copies.
static __global__ void msm_tree_plan_kernel(const uint32_t* __restrict__ hist, uint32_t* __restrict__ cursor,
                                            uint32_t* __restrict__ plan, int rounds) {
  __shared__ uint32_t G[TREE_MAX + 2];
  __shared__ uint32_t total;
  if (threadIdx.x == 0) {
    uint32_t run = 0;
    G[TREE_MAX + 1] = 0;
    for (int m = (int)TREE_MAX; m >= 1; m--) {
      run += hist[m];
      G[m] = run;
    }
    G[0] = run;
  }
  __syncthreads();
  for (uint32_t m = threadIdx.x; m <= TREE_MAX; m += blockDim.x) cursor[m] = m ? G[m + 1] : 0;
  auto g_at = [&](uint32_t x) -> uint32_t { return x <= TREE_MAX ? G[x] : 0u; };
  const uint32_t j = threadIdx.x;  // blockDim.x == TREE_JMAX
  for (int r = 0; r < rounds; r++) {
    uint32_t* rec = plan + (size_t)r * TREE_REC;
    uint32_t add = g_at(((2 * j + 1) << r) + 1);
    uint32_t all = (j == 0 && r > 0) ? add : g_at(((2 * j) << r) + 1);
    uint32_t ex = block_exclusive_scan(add, &total);
    rec[4 + j] = ex;
    if (j == 0) {
      rec[0] = total;
      rec[2] = G[1];  // light buckets (finish kernel)
      rec[4 + TREE_JMAX] = total;
    }
    __syncthreads();
    ex = block_exclusive_scan(all - add, &total);
    rec[4 + (TREE_JMAX + 1) + j] = ex;
    if (j == 0) {
      rec[1] = total;
      rec[4 + (TREE_JMAX + 1) + TREE_JMAX] = total;
    }
    __syncthreads();
  }
}

// ---- plan, step 3: buckets in descending order of size
static __global__ void msm_tree_emit_kernel(const uint32_t* __restrict__ offsets, uint32_t nbuckets,
                                            uint32_t* __restrict__ cursor, uint2* __restrict__ order) {
  __shared__ uint32_t h[TREE_MAX + 1];
  __shared__ uint32_t base[TREE_MAX + 1];
  for (uint32_t k = threadIdx.x; k <= TREE_MAX; k += blockDim.x) h[k] = 0;
  __syncthreads();
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t beg = 0, cnt = 0, local = 0;
  if (b < nbuckets) {
    beg = offsets[b];
    cnt = offsets[b + 1] - beg;
    if (cnt > TREE_MAX) cnt = 0;
    if (cnt) local = atomicAdd(&h[cnt], 1u);
  }
  __syncthreads();
  for (uint32_t k = threadIdx.x; k <= TREE_MAX; k += blockDim.x) base[k] = h[k] ? atomicAdd(&cursor[k], h[k]) : 0;
  __syncthreads();
  if (cnt) order[base[cnt] + local] = make_uint2(b | (cnt << TREE_BUCKET_BITS), beg);
}

// ---- rounds
template <class F>
struct TreeCfg {
  // small blocks, many of them: while one block waits for its inversion (one thread, ~60 us) the others on the SM
  // keep the integer pipe busy.  110 registers (G1) / 248 (G2).
  static constexpr int TB = 64;
  static constexpr int MIN_BLOCKS = sizeof(F) == 32 ? 8 : 4;
};

// products of the tree: the shared out-of-line bodies (the loop bodies stay small), dedicated squaring for Fp
template <class F>
ZKP_DEVINL F tree_mul(const F& a, const F& b) { return a * b; }
template <class F>
ZKP_DEVINL F tree_sqr(const F& a) { return a.sqr(); }
template <>
ZKP_DEVINL FpC tree_sqr<FpC>(const FpC& a) {
  Fp pa;
#pragma unroll
  for (int i = 0; i < 8; i++) pa.v[i] = a.v[i];
  Fp pr = mont_sqr_outlined<FpParams>(pa);
  FpC r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = pr.v[i];
  return r;
}

// Which items a block owns.  Items [0, T) of a round are cut into five segments: the first 3/4 in blocks of
// Bmax * {8,2,6,4,7,3,5,1} / 8 items per thread (a cycle of eight sizes: blocks that start together reach their
// inversions at different times), then 1/8, 1/16, 1/32, 1/32 of the items in blocks of Bmax / 2, / 4, / 8, / 16 items
// per thread: the blocks launched last are short, so the round does not end on a few long ones.
struct TreeSched {
  uint64_t c[6];
  uint32_t pref[6];
};
__host__ __device__ inline uint32_t tree_seg_items(uint32_t Bmax, int s) {
  uint32_t b = Bmax >> s;
  return b < 2 ? 2 : b;
}
__host__ __device__ inline uint32_t tree_sched(uint64_t T, uint32_t Bmax, uint32_t TB, TreeSched& sc) {
  sc.c[0] = 0;
  if (Bmax < 8) {
    for (int s = 1; s <= 5; s++) sc.c[s] = T;
  } else {
    sc.c[1] = T - (T >> 2);
    sc.c[2] = T - (T >> 3);
    sc.c[3] = T - (T >> 4);
    sc.c[4] = T - (T >> 5);
    sc.c[5] = T;
  }
  sc.pref[0] = 0;
  if (Bmax < 8) {
    sc.pref[1] = (uint32_t)((sc.c[1] + (uint64_t)TB * Bmax - 1) / ((uint64_t)TB * Bmax));
  } else {
    const uint64_t cycle = (uint64_t)TB * (Bmax / 8) * 36;
    sc.pref[1] = 8u * (uint32_t)((sc.c[1] + cycle - 1) / cycle);
  }
  for (int s = 1; s < 5; s++) {
    const uint64_t per = (uint64_t)TB * tree_seg_items(Bmax, s);
    sc.pref[s + 1] = sc.pref[s] + (uint32_t)((sc.c[s + 1] - sc.c[s] + per - 1) / per);
  }
  return sc.pref[5];
}
// block -> [blk0, end), items per thread; false when the block has nothing to do
__device__ inline bool tree_block_range(uint64_t T, uint32_t Bmax, uint32_t TB, uint32_t b, uint64_t& blk0, uint64_t& end,
                                        uint32_t& nB) {
  TreeSched sc;
  if (b >= tree_sched(T, Bmax, TB, sc)) return false;
  int s = 0;
  while (s < 4 && b >= sc.pref[s + 1]) s++;
  if (s == 0 && Bmax >= 8) {
    const uint32_t q = b >> 3, i = b & 7;
    const uint32_t PAT = 0x15374628u;  // nibble i = size of block i of a cycle, in units of Bmax / 8
    uint32_t before = 0;
    for (uint32_t k = 0; k < i; k++) before += (PAT >> (4 * k)) & 15u;
    const uint64_t unit = (uint64_t)TB * (Bmax / 8);
    nB = (Bmax / 8) * ((PAT >> (4 * i)) & 15u);
    blk0 = (uint64_t)q * unit * 36 + unit * before;
  } else {
    nB = s == 0 ? Bmax : tree_seg_items(Bmax, s);
    blk0 = sc.c[s] + (uint64_t)(b - sc.pref[s]) * TB * nB;
  }
  end = blk0 + (uint64_t)TB * nB;
  if (end > sc.c[s + 1]) end = sc.c[s + 1];
  return blk0 < end;
}

// smallest row whose items reach past id: start[] has TREE_JMAX + 1 entries, start[0] = 0, start[TREE_JMAX] > id
ZKP_DEVINL uint32_t tree_row(const uint32_t* start, uint32_t id) {
  uint32_t lo = 0, hi = TREE_JMAX;
#pragma unroll 1
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (start[mid] <= id) lo = mid;
    else hi = mid;
  }
  return lo;
}

struct TreeItem {
  uint32_t b;      // bucket
  uint64_t in0;    // first operand (FIRST: index into `sorted`; later rounds: index into the input buffer)
  uint64_t out;    // slot in the output buffer
  bool last;       // this item leaves the bucket with one point: the result goes to `buckets`
};

ZKP_DEVINL uint64_t tree_base(uint2 ent, int r) {
  const uint32_t b = ent.x & ((1u << TREE_BUCKET_BITS) - 1);
  uint64_t base = ent.y;
  for (int i = 0; i < r; i++) base = (base + b) >> 1;
  return base;
}
ZKP_DEVINL TreeItem tree_item(uint2 ent, uint32_t j, int r) {
  TreeItem it;
  it.b = ent.x & ((1u << TREE_BUCKET_BITS) - 1);
  const uint32_t m0 = ent.x >> TREE_BUCKET_BITS;
  const uint64_t base = tree_base(ent, r);
  it.in0 = base + 2 * (uint64_t)j;
  it.out = ((base + it.b) >> 1) + j;
  const uint32_t mr = (m0 + (1u << r) - 1) >> r;
  it.last = mr <= 2;
  return it;
}

template <class F, bool FIRST>
ZKP_DEVINL Affine<F> tree_load(const Affine<F>* __restrict__ pts, const uint32_t* __restrict__ sorted,
                               const Affine<F>* __restrict__ inbuf, uint64_t idx) {
  if constexpr (FIRST) {
    uint32_t e = sorted[idx];
    Affine<F> p = pts[e & 0x7fffffffu];
    if (e >> 31) p.y = p.y.neg();
    return p;
  } else {
    return inbuf[idx];
  }
}
template <class F, bool FIRST>
ZKP_DEVINL F tree_load_x(const Affine<F>* __restrict__ pts, const uint32_t* __restrict__ sorted,
                         const Affine<F>* __restrict__ inbuf, uint64_t idx) {
  if constexpr (FIRST) return pts[sorted[idx] & 0x7fffffffu].x;
  else return inbuf[idx].x;
}

template <class F>
ZKP_DEVINL void tree_store(const TreeItem& it, const Affine<F>& res, Affine<F>* __restrict__ outbuf,
                           XYZZ<F>* __restrict__ buckets) {
  if (it.last) buckets[it.b] = XYZZ<F>::from_affine(res);
  else outbuf[it.out] = res;
}

// One round.  A block owns a run of items of the round (tree_block_range): additions first, then copies.
// Pass 1 (k ascending): thread t multiplies up the denominators of its additions blk0 + k*TB + t, leaving the
// running product before each one in `pre` (global memory, indexed by item: 32 bytes written and read per addition).
// The TB thread products are multiplied up a binary tree in shared memory, the root is inverted once (binary Euclid,
// one thread), the inverses come back down the tree.  Pass 2 (k descending): 1/d_k = inv * pre_k, inv *= d_k, chord
// rule.  The operands are loaded again in pass 2 (pass 1 only reads the x coordinates): holding them would cost
// 128 bytes of shared memory per addition.
template <class F, bool FIRST>
__global__ void __launch_bounds__(TreeCfg<F>::TB, TreeCfg<F>::MIN_BLOCKS)
msm_tree_round_kernel(const Affine<F>* __restrict__ pts, const uint32_t* __restrict__ sorted,
                      const Affine<F>* __restrict__ inbuf, Affine<F>* __restrict__ outbuf,
                      const uint2* __restrict__ order, const uint32_t* __restrict__ rec, int r, uint32_t Bmax,
                      F* __restrict__ pre, XYZZ<F>* __restrict__ buckets) {
  constexpr int TB = TreeCfg<F>::TB;
  __shared__ uint32_t s_add[TREE_JMAX + 1], s_copy[TREE_JMAX + 1];
  __shared__ F s_tree[2 * TB];
  const uint32_t NA = rec[0], NC = rec[1];
  uint64_t blk0, end;
  uint32_t nB;
  if (!tree_block_range((uint64_t)NA + NC, Bmax, TB, blockIdx.x, blk0, end, nB)) return;
  for (uint32_t k = threadIdx.x; k <= TREE_JMAX; k += TB) {
    s_add[k] = rec[4 + k];
    s_copy[k] = rec[4 + (TREE_JMAX + 1) + k];
  }
  __syncthreads();
  const uint32_t tid = threadIdx.x;

  F acc = F::one();
  uint32_t ja = 0, jc = 0;
  bool ja_set = false, jc_set = false;
#pragma unroll 1
  for (uint32_t k = 0; k < nB; k++) {
    const uint64_t id64 = blk0 + (uint64_t)k * TB + tid;
    if (id64 >= end) break;
    if (id64 < NA) {
      const uint32_t id = (uint32_t)id64;
      if (!ja_set) {
        ja = tree_row(s_add, id);
        ja_set = true;
      } else {
        while (s_add[ja + 1] <= id) ja++;
      }
      const TreeItem it = tree_item(order[id - s_add[ja]], ja, r);
      const F x1 = tree_load_x<F, FIRST>(pts, sorted, inbuf, it.in0);
      const F x2 = tree_load_x<F, FIRST>(pts, sorted, inbuf, it.in0 + 1);
      F d = x2 - x1;
      bool regular = true;
      if (x1.is_zero() || x2.is_zero() || d.is_zero()) {  // rare: infinity among the operands, P + P, P + (-P)
        const Affine<F> p = tree_load<F, FIRST>(pts, sorted, inbuf, it.in0);
        const Affine<F> q = tree_load<F, FIRST>(pts, sorted, inbuf, it.in0 + 1);
        if (p.is_inf() || q.is_inf()) regular = false;
        else if (d.is_zero()) {
          if (p.y == q.y && !p.y.is_zero()) d = p.y.dbl();
          else regular = false;
        }
      }
      if (regular) {
        pre[id64] = acc;
        acc = tree_mul(acc, d);
      }
    } else {  // odd last entry of a bucket: moves to the next round as it is
      const uint32_t cid = (uint32_t)(id64 - NA);
      if (!jc_set) {
        jc = tree_row(s_copy, cid);
        jc_set = true;
      } else {
        while (s_copy[jc + 1] <= cid) jc++;
      }
      const uint32_t rank = (s_add[jc + 1] - s_add[jc]) + (cid - s_copy[jc]);
      TreeItem it = tree_item(order[rank], jc, r);
      it.last = (jc == 0);  // m_r == 1 (only in the first round): the entry is the bucket
      tree_store<F>(it, tree_load<F, FIRST>(pts, sorted, inbuf, it.in0), outbuf, buckets);
    }
  }

  // ---- one inversion for the block: products up the tree, root inverted, inverses down
  s_tree[TB + tid] = acc;
  __syncthreads();
#pragma unroll 1
  for (int w = TB >> 1; w >= 1; w >>= 1) {
    if ((int)tid < w) s_tree[w + tid] = tree_mul(s_tree[2 * (w + tid)], s_tree[2 * (w + tid) + 1]);
    __syncthreads();
  }
  if (tid == 0) s_tree[1] = s_tree[1].inv();
  __syncthreads();
#pragma unroll 1
  for (int w = 1; w < TB; w <<= 1) {
    if ((int)tid < w) {
      const F up = s_tree[w + tid], l = s_tree[2 * (w + tid)], rr = s_tree[2 * (w + tid) + 1];
      s_tree[2 * (w + tid)] = tree_mul(up, rr);
      s_tree[2 * (w + tid) + 1] = tree_mul(up, l);
    }
    __syncthreads();
  }
  F inv = s_tree[TB + tid];

#pragma unroll 1
  for (int k = (int)nB - 1; k >= 0; k--) {
    const uint64_t id64 = blk0 + (uint64_t)k * TB + tid;
    if (id64 >= end || id64 >= NA) continue;
    const uint32_t id = (uint32_t)id64;
    while (s_add[ja] > id) ja--;  // ja was left at the row of this thread's last addition
    const TreeItem it = tree_item(order[id - s_add[ja]], ja, r);
    const Affine<F> p = tree_load<F, FIRST>(pts, sorted, inbuf, it.in0);
    const Affine<F> q = tree_load<F, FIRST>(pts, sorted, inbuf, it.in0 + 1);
    Affine<F> res;
    if (p.is_inf()) res = q;
    else if (q.is_inf()) res = p;
    else {
      F d = q.x - p.x, num = q.y - p.y;
      bool regular = true;
      if (d.is_zero()) {
        if (num.is_zero() && !p.y.is_zero()) {  // P + P: tangent
          d = p.y.dbl();
          const F xx = tree_sqr(p.x);
          num = xx.dbl() + xx;
        } else {
          regular = false;
          res = Affine<F>::inf();
        }
      }
      if (regular) {
        const F dinv = tree_mul(inv, pre[id64]);
        inv = tree_mul(inv, d);
        const F lam = tree_mul(num, dinv);
        res.x = tree_sqr(lam) - p.x - q.x;
        res.y = tree_mul(lam, p.x - res.x) - p.y;
      }
    }
    tree_store<F>(it, res, outbuf, buckets);
  }
}

// After K rounds: the m_K = ceil(m_0 / 2^K) points a bucket still holds, as one XYZZ chain (mixed additions, the
// accumulate kernel's own loop body).  One thread per bucket in descending order of size; a bucket with m_K == 1
// was finished by the round that left it with one point.
template <class F, int OUTLINE>
__global__ void __launch_bounds__(128, AccMinBlocks<F>::value)
msm_tree_finish_kernel(const Affine<F>* __restrict__ inbuf, const uint2* __restrict__ order,
                       const uint32_t* __restrict__ rec, int K, XYZZ<F>* __restrict__ buckets) {
  const uint32_t rank = blockIdx.x * blockDim.x + threadIdx.x;
  if (rank >= rec[2]) return;
  const uint2 ent = order[rank];
  const uint32_t m0 = ent.x >> TREE_BUCKET_BITS;
  const uint32_t mK = (m0 + (1u << K) - 1) >> K;
  if (mK < 2) return;
  const Affine<F>* p = inbuf + tree_base(ent, K);
  XYZZ<F> acc = XYZZ<F>::from_affine(p[0]);
  for (uint32_t i = 1; i < mK; i++) acc_madd<F, OUTLINE>(acc, p[i]);
  buckets[ent.x & ((1u << TREE_BUCKET_BITS) - 1)] = acc;
}

}  // namespace zkp
