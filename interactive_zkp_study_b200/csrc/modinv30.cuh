// Modular inversion by batched division steps (Bernstein-Yang "safegcd", in the 30-bit-limb form popularised by
// libsecp256k1's modinv32): x^-1 mod MOD for an odd 254-bit modulus, on one thread, with integer arithmetic only.
//
// Used where the reference calls py_ecc's prime_field_inv through an affine addition or a normalisation: once per
// MSM result (/root/reference/zkp/plonk/kzg.py:59-67 returns an affine point), in the batched conversions and in
// the Fr batch inverse.  The binary extended Euclid it replaces (Mont256::inv_euclid, kept as a cross-check) walks
// the operands one bit at a time with four 256-bit values in flight (~20 000 dependent instructions, 57 us on a
// lone B200 thread); here 30 division steps run on single 32-bit registers and their combined effect -- a 2x2
// matrix of 31-bit integers -- is applied to the 270-bit values once per 30 steps: ~19 rounds of (30 short steps +
// 2 x 9 limb rows of multiply-adds).
//
// State: f (odd, starts as MOD), g (starts as x), and d, e with d * x == f, e * x == g (mod MOD) throughout.
// One division step with eta = -delta:
//     g odd and eta < 0 :  (eta, f, g) <- (-eta - 1, g, (g - f) / 2)
//     otherwise         :  (eta, f, g) <- (eta - 1, f, (g + (g & 1) f) / 2)
// After enough steps g == 0 and f == +-gcd = +-1, so x^-1 == +-d.
//
// The file is plain C++ apart from the ZKP_HD qualifier, so the same source is compiled by g++ in the CPU test
// (tests/test_cpu_modinv.py) and checked there against Python's pow(x, -1, MOD).
#pragma once
#include <cstdint>

#ifndef ZKP_HD
#ifdef __CUDACC__
#define ZKP_HD __host__ __device__ __forceinline__
#define ZKP_UNROLL _Pragma("unroll")
#define ZKP_NOUNROLL _Pragma("unroll 1")
#else
#define ZKP_HD inline
#define ZKP_UNROLL
#define ZKP_NOUNROLL
#endif
#endif

namespace zkp {
namespace modinv30 {

static constexpr int32_t M30 = (int32_t)(0xffffffffu >> 2);

struct S30 {
  int32_t v[9];  // value = sum v[i] 2^(30 i); limbs 0..7 in [0, 2^30), limb 8 signed
};
struct Trans {
  int32_t u, v, q, r;
};

// limb i (30 bits) of the 256-bit little-endian integer held in 8 x 32-bit words
template <class W>
ZKP_HD constexpr int32_t limb30(const W& word, int i) {
  const int bit = 30 * i, w = bit >> 5, sh = bit & 31;
  uint64_t two = (uint64_t)word(w) | (w + 1 < 8 ? ((uint64_t)word(w + 1) << 32) : 0ull);
  return (int32_t)((two >> sh) & (uint64_t)M30);
}
// MOD^-1 mod 2^30 (MOD odd): Newton steps double the number of correct low bits
ZKP_HD constexpr uint32_t inv30_of(uint32_t m0) {
  uint32_t x = m0;                    // correct to 3 bits (odd m: m * m == 1 mod 8)
  for (int i = 0; i < 5; i++) x *= 2u - m0 * x;
  return x & (uint32_t)M30;
}

// 30 division steps on the low 30 bits of f and g; returns the new eta and the transition matrix t with
// 2^30 [f', g'] = t [f, g].  Branch free (the CPU build and every lane of a warp run the same instructions).
ZKP_HD int32_t divsteps_30(int32_t eta, uint32_t f0, uint32_t g0, Trans& t) {
  uint32_t u = 1, v = 0, q = 0, r = 1;
  uint32_t f = f0, g = g0;
  ZKP_UNROLL
  for (int i = 0; i < 30; i++) {
    uint32_t c1 = (uint32_t)(eta >> 31);      // all ones when eta < 0
    const uint32_t c2 = 0u - (g & 1u);        // all ones when g is odd
    const uint32_t x = (f ^ c1) - c1, y = (u ^ c1) - c1, z = (v ^ c1) - c1;  // (f, u, v) negated when eta < 0
    g += x & c2;
    q += y & c2;
    r += z & c2;
    c1 &= c2;                                  // both: the roles of f and g swap
    eta = (int32_t)(((uint32_t)eta ^ c1) - (c1 + 1u));
    f += g & c1;
    u += q & c1;
    v += r & c1;
    g >>= 1;
    u <<= 1;
    v <<= 1;
  }
  t.u = (int32_t)u;
  t.v = (int32_t)v;
  t.q = (int32_t)q;
  t.r = (int32_t)r;
  return eta;
}

// [d, e] <- t [d, e] / 2^30 (mod MOD), both kept in (-2 MOD, MOD)
template <class P>
ZKP_HD void update_de(S30& d, S30& e, const Trans& t) {
  auto word = [](int i) { return P::MOD_(i); };
  const uint32_t minv = inv30_of(P::MOD_(0));
  const int32_t u = t.u, v = t.v, q = t.q, r = t.r;
  const int32_t sd = d.v[8] >> 31, se = e.v[8] >> 31;
  int32_t md = (u & sd) + (v & se);  // multiples of MOD added so that the results stay in range
  int32_t me = (q & sd) + (r & se);
  int32_t di = d.v[0], ei = e.v[0];
  int64_t cd = (int64_t)u * di + (int64_t)v * ei;
  int64_t ce = (int64_t)q * di + (int64_t)r * ei;
  // ... and adjusted so that the low 30 bits of t [d, e] + MOD [md, me] vanish
  md -= (int32_t)((minv * (uint32_t)cd + (uint32_t)md) & (uint32_t)M30);
  me -= (int32_t)((minv * (uint32_t)ce + (uint32_t)me) & (uint32_t)M30);
  cd += (int64_t)limb30(word, 0) * md;
  ce += (int64_t)limb30(word, 0) * me;
  cd >>= 30;
  ce >>= 30;
  ZKP_UNROLL
  for (int i = 1; i < 9; i++) {
    di = d.v[i];
    ei = e.v[i];
    cd += (int64_t)u * di + (int64_t)v * ei;
    ce += (int64_t)q * di + (int64_t)r * ei;
    cd += (int64_t)limb30(word, i) * md;
    ce += (int64_t)limb30(word, i) * me;
    d.v[i - 1] = (int32_t)cd & M30;
    cd >>= 30;
    e.v[i - 1] = (int32_t)ce & M30;
    ce >>= 30;
  }
  d.v[8] = (int32_t)cd;
  e.v[8] = (int32_t)ce;
}

// [f, g] <- t [f, g] / 2^30 (exact)
ZKP_HD void update_fg(S30& f, S30& g, const Trans& t) {
  const int32_t u = t.u, v = t.v, q = t.q, r = t.r;
  int32_t fi = f.v[0], gi = g.v[0];
  int64_t cf = (int64_t)u * fi + (int64_t)v * gi;
  int64_t cg = (int64_t)q * fi + (int64_t)r * gi;
  cf >>= 30;
  cg >>= 30;
  ZKP_UNROLL
  for (int i = 1; i < 9; i++) {
    fi = f.v[i];
    gi = g.v[i];
    cf += (int64_t)u * fi + (int64_t)v * gi;
    cg += (int64_t)q * fi + (int64_t)r * gi;
    f.v[i - 1] = (int32_t)cf & M30;
    cf >>= 30;
    g.v[i - 1] = (int32_t)cg & M30;
    cg >>= 30;
  }
  f.v[8] = (int32_t)cf;
  g.v[8] = (int32_t)cg;
}

// r in (-2 MOD, MOD) -> [0, MOD), negated first when sign < 0
template <class P>
ZKP_HD void normalize(S30& r, int32_t sign) {
  auto word = [](int i) { return P::MOD_(i); };
  int32_t cond_add = r.v[8] >> 31;
  ZKP_UNROLL
  for (int i = 0; i < 9; i++) r.v[i] += limb30(word, i) & cond_add;
  const int32_t cond_negate = sign >> 31;
  ZKP_UNROLL
  for (int i = 0; i < 9; i++) r.v[i] = (r.v[i] ^ cond_negate) - cond_negate;
  ZKP_UNROLL
  for (int i = 0; i < 8; i++) {
    r.v[i + 1] += r.v[i] >> 30;
    r.v[i] &= M30;
  }
  cond_add = r.v[8] >> 31;
  ZKP_UNROLL
  for (int i = 0; i < 9; i++) r.v[i] += limb30(word, i) & cond_add;
  ZKP_UNROLL
  for (int i = 0; i < 8; i++) {
    r.v[i + 1] += r.v[i] >> 30;
    r.v[i] &= M30;
  }
}

// out = x^-1 mod MOD as 8 x 32-bit words (x in [0, MOD), 0 -> 0)
template <class P>
ZKP_HD void inverse(const uint32_t (&x)[8], uint32_t (&out)[8]) {
  auto mword = [](int i) { return P::MOD_(i); };
  S30 d, e, f, g;
  ZKP_UNROLL
  for (int i = 0; i < 9; i++) {
    const int bit = 30 * i, w = bit >> 5, sh = bit & 31;
    uint64_t two = (uint64_t)x[w] | (w + 1 < 8 ? ((uint64_t)x[w + 1] << 32) : 0ull);
    g.v[i] = (int32_t)((two >> sh) & (uint64_t)M30);
    f.v[i] = limb30(mword, i);
    d.v[i] = 0;
    e.v[i] = i == 0 ? 1 : 0;
  }
  int32_t eta = -1;
  // at most 741 steps are needed for 256-bit operands (Bernstein-Yang, delta starting at 1): 25 rounds
  ZKP_NOUNROLL
  for (int round = 0; round < 25; round++) {
    Trans t;
    eta = divsteps_30(eta, (uint32_t)f.v[0] | ((uint32_t)f.v[1] << 30), (uint32_t)g.v[0] | ((uint32_t)g.v[1] << 30), t);
    update_de<P>(d, e, t);
    update_fg(f, g, t);
    int32_t nz = 0;
    ZKP_UNROLL
    for (int i = 0; i < 9; i++) nz |= g.v[i];
    if (nz == 0) break;
  }
  normalize<P>(d, f.v[8]);  // f == -1: the inverse is -d
  // x == 0: g is 0 from the start, f = MOD, d = 0 -> 0, as py_ecc's prime_field_inv(0)
  ZKP_UNROLL
  for (int w = 0; w < 8; w++) {
    const int bit = 32 * w, i = bit / 30, sh = bit % 30;
    uint64_t acc = (uint64_t)(uint32_t)d.v[i] >> sh;
    acc |= (uint64_t)(uint32_t)d.v[i + 1] << (30 - sh);
    if (i + 2 < 9) acc |= (uint64_t)(uint32_t)d.v[i + 2] << (60 - sh);
    out[w] = (uint32_t)acc;
  }
}

}  // namespace modinv30
}  // namespace zkp
