// Montgomery arithmetic for the two 254-bit BN254 prime fields (Fp, Fr) on sm_100a.
//
// Replaces, on the device, the per-element Python arithmetic of py_ecc's FQ as the
// reference uses it (/root/reference/zkp/plonk/field.py:36-51 `FR(FQ)`,
// /root/reference/zkp/groth16/poly_utils.py:12-13, and every FQ op inside
// bn128.add/double/multiply, call sites /root/reference/zkp/groth16/proving.py:12-15).
//
// Representation: 8 x 32-bit little-endian limbs, Montgomery radix 2^256, always fully
// reduced to [0, MOD) so equality is limb equality.  The multiplier is a row-wise CIOS
// built from 32x32+64 wide MADs (PTX mad.lo.cc/madc.hi.cc pairs, which ptxas fuses into
// IMAD.WIDE.U32[.X] with a predicate carry).  Products of even and odd limbs are kept in
// two accumulators whose 64-bit lanes never overlap, so each half-row is one 4-deep carry
// chain and the two half-rows are independent (ILP 2):
//
//     T = X + Y * 2^32,  X lanes at limb pairs (0,1)(2,3)(4,5)(6,7), Y at (1,2)...(7,8)
//
// Per row: X += a_even*b_i, Y += a_odd*b_i, q = X[0]*N0INV, X += n_even*q, Y += n_odd*q;
// then X[0] == 0, the roles of X and Y swap (a one-limb shift of T) and the surviving
// limb X[1] is carried into the next row's first add.  128 IMAD.WIDE + 8 IMAD per
// multiplication, no 32-bit carry propagation across the whole accumulator inside the loop.
#pragma once
#include <cstdint>
#include "bn254_params.cuh"
#include "modinv30.cuh"

namespace zkp {

#define ZKP_DEVINL __device__ __forceinline__

namespace detail {

// acc (8 limbs, four 64-bit lanes) += {m0,m1,m2,m3} * b, lane k gets m_k*b; returns the carry out.
ZKP_DEVINL uint32_t mad_lanes(uint32_t (&acc)[8], uint32_t m0, uint32_t m1, uint32_t m2, uint32_t m3,
                              uint32_t b) {
  uint32_t c;
  asm("mad.lo.cc.u32   %0, %9,  %13, %0;\n\t"
      "madc.hi.cc.u32  %1, %9,  %13, %1;\n\t"
      "madc.lo.cc.u32  %2, %10, %13, %2;\n\t"
      "madc.hi.cc.u32  %3, %10, %13, %3;\n\t"
      "madc.lo.cc.u32  %4, %11, %13, %4;\n\t"
      "madc.hi.cc.u32  %5, %11, %13, %5;\n\t"
      "madc.lo.cc.u32  %6, %12, %13, %6;\n\t"
      "madc.hi.cc.u32  %7, %12, %13, %7;\n\t"
      "addc.u32        %8, 0, 0;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]),
        "+r"(acc[6]), "+r"(acc[7]), "=r"(c)
      : "r"(m0), "r"(m1), "r"(m2), "r"(m3), "r"(b));
  return c;
}

// The same with the carry out added straight into `top` (the limb above the lanes, held by the other accumulator):
// one IADD3.X.  Returned as a value and added by the caller (`Y[7] += c`) ptxas made it VIADD + predicated IMAD.MOV +
// IMAD.MOV.U32 -- two of them on the multiplier pipe, 32 per product.
ZKP_DEVINL void mad_lanes_top(uint32_t (&acc)[8], uint32_t m0, uint32_t m1, uint32_t m2, uint32_t m3,
                                  uint32_t b, uint32_t& top) {
  asm("mad.lo.cc.u32   %0, %9,  %13, %0;\n\t"
      "madc.hi.cc.u32  %1, %9,  %13, %1;\n\t"
      "madc.lo.cc.u32  %2, %10, %13, %2;\n\t"
      "madc.hi.cc.u32  %3, %10, %13, %3;\n\t"
      "madc.lo.cc.u32  %4, %11, %13, %4;\n\t"
      "madc.hi.cc.u32  %5, %11, %13, %5;\n\t"
      "madc.lo.cc.u32  %6, %12, %13, %6;\n\t"
      "madc.hi.cc.u32  %7, %12, %13, %7;\n\t"
      "addc.u32        %8, %8, 0;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]),
        "+r"(acc[6]), "+r"(acc[7]), "+r"(top)
      : "r"(m0), "r"(m1), "r"(m2), "r"(m3), "r"(b));
}

// x0 += left (one limb below the first lane of acc); the carry of that add enters acc's
// first lane, then acc += {m0..m3} * b.  The top lane cannot carry out (see ff.cuh header:
// Y * 2^32 <= T < 2^288).
ZKP_DEVINL void mad_lanes_cin(uint32_t& x0, uint32_t left, uint32_t (&acc)[8], uint32_t m0, uint32_t m1,
                              uint32_t m2, uint32_t m3, uint32_t b) {
  asm("add.cc.u32      %8, %8, %9;\n\t"
      "madc.lo.cc.u32  %0, %10, %14, %0;\n\t"
      "madc.hi.cc.u32  %1, %10, %14, %1;\n\t"
      "madc.lo.cc.u32  %2, %11, %14, %2;\n\t"
      "madc.hi.cc.u32  %3, %11, %14, %3;\n\t"
      "madc.lo.cc.u32  %4, %12, %14, %4;\n\t"
      "madc.hi.cc.u32  %5, %12, %14, %5;\n\t"
      "madc.lo.cc.u32  %6, %13, %14, %6;\n\t"
      "madc.hi.u32     %7, %13, %14, %7;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]),
        "+r"(acc[6]), "+r"(acc[7]), "+r"(x0)
      : "r"(left), "r"(m0), "r"(m1), "r"(m2), "r"(m3), "r"(b));
}

// x0 += left; the carry ripples through acc (no product): the row opening of a pure reduction row
ZKP_DEVINL void add_cin(uint32_t& x0, uint32_t left, uint32_t (&acc)[8]) {
  asm("add.cc.u32   %8, %8, %9;\n\t"
      "addc.cc.u32  %0, %0, 0;\n\t"
      "addc.cc.u32  %1, %1, 0;\n\t"
      "addc.cc.u32  %2, %2, 0;\n\t"
      "addc.cc.u32  %3, %3, 0;\n\t"
      "addc.cc.u32  %4, %4, 0;\n\t"
      "addc.cc.u32  %5, %5, 0;\n\t"
      "addc.cc.u32  %6, %6, 0;\n\t"
      "addc.u32     %7, %7, 0;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]),
        "+r"(x0)
      : "r"(left));
}

// ---- partial chains for the dedicated squaring (sqr_inline): the same lane accumulators, but only lanes
// K..3 receive a product (the lower multiplicand limbs of a squaring row are absent), generated text.
template <int K>
ZKP_DEVINL uint32_t mad_lanes_from(uint32_t (&acc)[8], uint32_t m0, uint32_t m1, uint32_t m2, uint32_t m3, uint32_t b) {
  uint32_t c = 0;
  if constexpr (K == 0) {
    asm(
        "mad.lo.cc.u32   %0, %9, %13, %0;\n\t"
        "madc.hi.cc.u32  %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32  %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32  %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32  %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32  %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32  %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32  %7, %12, %13, %7;\n\t"
        "addc.u32        %8, 0, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]),
          "+r"(acc[6]), "+r"(acc[7]), "=r"(c)
        : "r"(m0), "r"(m1), "r"(m2), "r"(m3), "r"(b));
  }
  else if constexpr (K == 1) {
    asm(
        "mad.lo.cc.u32   %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32  %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32  %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32  %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32  %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32  %7, %12, %13, %7;\n\t"
        "addc.u32        %8, 0, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]),
          "+r"(acc[6]), "+r"(acc[7]), "=r"(c)
        : "r"(m0), "r"(m1), "r"(m2), "r"(m3), "r"(b));
  }
  else if constexpr (K == 2) {
    asm(
        "mad.lo.cc.u32   %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32  %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32  %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32  %7, %12, %13, %7;\n\t"
        "addc.u32        %8, 0, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]),
          "+r"(acc[6]), "+r"(acc[7]), "=r"(c)
        : "r"(m0), "r"(m1), "r"(m2), "r"(m3), "r"(b));
  }
  else if constexpr (K == 3) {
    asm(
        "mad.lo.cc.u32   %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32  %7, %12, %13, %7;\n\t"
        "addc.u32        %8, 0, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]),
          "+r"(acc[6]), "+r"(acc[7]), "=r"(c)
        : "r"(m0), "r"(m1), "r"(m2), "r"(m3), "r"(b));
  }
  return c;  // K == 4: no lane takes a product
}

template <int K>
ZKP_DEVINL void mad_lanes_from_top(uint32_t (&acc)[8], uint32_t m0, uint32_t m1, uint32_t m2, uint32_t m3, uint32_t b,
                                   uint32_t& top) {
  if constexpr (K == 0) {
    asm(
        "mad.lo.cc.u32   %0, %9, %13, %0;\n\t"
        "madc.hi.cc.u32  %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32  %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32  %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32  %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32  %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32  %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32  %7, %12, %13, %7;\n\t"
        "addc.u32        %8, %8, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]),
          "+r"(acc[6]), "+r"(acc[7]), "+r"(top)
        : "r"(m0), "r"(m1), "r"(m2), "r"(m3), "r"(b));
  }
  else if constexpr (K == 1) {
    asm(
        "mad.lo.cc.u32   %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32  %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32  %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32  %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32  %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32  %7, %12, %13, %7;\n\t"
        "addc.u32        %8, %8, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]),
          "+r"(acc[6]), "+r"(acc[7]), "+r"(top)
        : "r"(m0), "r"(m1), "r"(m2), "r"(m3), "r"(b));
  }
  else if constexpr (K == 2) {
    asm(
        "mad.lo.cc.u32   %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32  %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32  %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32  %7, %12, %13, %7;\n\t"
        "addc.u32        %8, %8, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]),
          "+r"(acc[6]), "+r"(acc[7]), "+r"(top)
        : "r"(m0), "r"(m1), "r"(m2), "r"(m3), "r"(b));
  }
  else if constexpr (K == 3) {
    asm(
        "mad.lo.cc.u32   %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32  %7, %12, %13, %7;\n\t"
        "addc.u32        %8, %8, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]),
          "+r"(acc[6]), "+r"(acc[7]), "+r"(top)
        : "r"(m0), "r"(m1), "r"(m2), "r"(m3), "r"(b));
  }
  // K == 4: no lane takes a product
}


// x0 += left, the carry ripples through the lanes below K and enters the product chain of lanes K..3
template <int K>
ZKP_DEVINL void mad_lanes_cin_from(uint32_t& x0, uint32_t left, uint32_t (&acc)[8], uint32_t m0, uint32_t m1, uint32_t m2,
                                   uint32_t m3, uint32_t b) {
  if constexpr (K == 0) {
    asm(
        "add.cc.u32      %8, %8, %9;\n\t"
        "madc.lo.cc.u32  %0, %10, %14, %0;\n\t"
        "madc.hi.cc.u32  %1, %10, %14, %1;\n\t"
        "madc.lo.cc.u32  %2, %11, %14, %2;\n\t"
        "madc.hi.cc.u32  %3, %11, %14, %3;\n\t"
        "madc.lo.cc.u32  %4, %12, %14, %4;\n\t"
        "madc.hi.cc.u32  %5, %12, %14, %5;\n\t"
        "madc.lo.cc.u32  %6, %13, %14, %6;\n\t"
        "madc.hi.u32     %7, %13, %14, %7;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]),
          "+r"(acc[6]), "+r"(acc[7]), "+r"(x0)
        : "r"(left), "r"(m0), "r"(m1), "r"(m2), "r"(m3), "r"(b));
  }
  else if constexpr (K == 1) {
    asm(
        "add.cc.u32      %8, %8, %9;\n\t"
        "addc.cc.u32     %0, %0, 0;\n\t"
        "addc.cc.u32     %1, %1, 0;\n\t"
        "madc.lo.cc.u32  %2, %11, %14, %2;\n\t"
        "madc.hi.cc.u32  %3, %11, %14, %3;\n\t"
        "madc.lo.cc.u32  %4, %12, %14, %4;\n\t"
        "madc.hi.cc.u32  %5, %12, %14, %5;\n\t"
        "madc.lo.cc.u32  %6, %13, %14, %6;\n\t"
        "madc.hi.u32     %7, %13, %14, %7;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]),
          "+r"(acc[6]), "+r"(acc[7]), "+r"(x0)
        : "r"(left), "r"(m0), "r"(m1), "r"(m2), "r"(m3), "r"(b));
  }
  else if constexpr (K == 2) {
    asm(
        "add.cc.u32      %8, %8, %9;\n\t"
        "addc.cc.u32     %0, %0, 0;\n\t"
        "addc.cc.u32     %1, %1, 0;\n\t"
        "addc.cc.u32     %2, %2, 0;\n\t"
        "addc.cc.u32     %3, %3, 0;\n\t"
        "madc.lo.cc.u32  %4, %12, %14, %4;\n\t"
        "madc.hi.cc.u32  %5, %12, %14, %5;\n\t"
        "madc.lo.cc.u32  %6, %13, %14, %6;\n\t"
        "madc.hi.u32     %7, %13, %14, %7;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]),
          "+r"(acc[6]), "+r"(acc[7]), "+r"(x0)
        : "r"(left), "r"(m0), "r"(m1), "r"(m2), "r"(m3), "r"(b));
  }
  else if constexpr (K == 3) {
    asm(
        "add.cc.u32      %8, %8, %9;\n\t"
        "addc.cc.u32     %0, %0, 0;\n\t"
        "addc.cc.u32     %1, %1, 0;\n\t"
        "addc.cc.u32     %2, %2, 0;\n\t"
        "addc.cc.u32     %3, %3, 0;\n\t"
        "addc.cc.u32     %4, %4, 0;\n\t"
        "addc.cc.u32     %5, %5, 0;\n\t"
        "madc.lo.cc.u32  %6, %13, %14, %6;\n\t"
        "madc.hi.u32     %7, %13, %14, %7;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]),
          "+r"(acc[6]), "+r"(acc[7]), "+r"(x0)
        : "r"(left), "r"(m0), "r"(m1), "r"(m2), "r"(m3), "r"(b));
  }
}

// One CIOS row.  X: lanes aligned with the current limb 0; Y: lanes one limb higher.
template <class P>
ZKP_DEVINL void mont_row(uint32_t (&X)[8], uint32_t (&Y)[8], uint32_t left, const uint32_t (&a)[8],
                         uint32_t bi) {
  mad_lanes_cin(X[0], left, Y, a[1], a[3], a[5], a[7], bi);
  mad_lanes_top(X, a[0], a[2], a[4], a[6], bi, Y[7]);
  uint32_t q = X[0] * P::N0INV;
  mad_lanes_top(X, P::MOD_(0), P::MOD_(2), P::MOD_(4), P::MOD_(6), q, Y[7]);  // Y[7] + carry <= its final value: no overflow
  uint32_t cy = mad_lanes(Y, P::MOD_(1), P::MOD_(3), P::MOD_(5), P::MOD_(7), q);
  (void)cy;  // provably 0
}

// One row of the dedicated squaring: row I multiplies a_I by the limbs {a_I, 2 a_{I+1}, ...} of
// a_I 2^(32 I) + 2 (a >> 32 (I + 1)) 2^(32 (I + 1)) -- each cross product a_i a_j is taken once, doubled --
// so the multiplicand limbs below I are absent and their lanes only pass the carry on.
template <class P, int I>
ZKP_DEVINL void mont_row_sq(uint32_t (&X)[8], uint32_t (&Y)[8], uint32_t left, const uint32_t (&m)[8], uint32_t bi) {
  constexpr int KX = (I + 1) / 2, KY = I / 2;  // first lane of X (even limbs) / Y (odd limbs) that takes a product
  mad_lanes_cin_from<KY>(X[0], left, Y, m[1], m[3], m[5], m[7], bi);
  mad_lanes_from_top<KX>(X, m[0], m[2], m[4], m[6], bi, Y[7]);
  uint32_t q = X[0] * P::N0INV;
  mad_lanes_top(X, P::MOD_(0), P::MOD_(2), P::MOD_(4), P::MOD_(6), q, Y[7]);
  uint32_t cy = mad_lanes(Y, P::MOD_(1), P::MOD_(3), P::MOD_(5), P::MOD_(7), q);
  (void)cy;
}

}  // namespace detail

template <class P, bool COMPACT>
struct Mont256;

// Out-of-line multiplication (arguments and result travel in registers): used by the COMPACT
// flavour of the field type.  The serial phases of the MSM (bucket reduction, Horner, inversion) run
// a handful of warps; with every product inlined their code is >100 KB and each lone warp stalls on
// instruction fetch, while with one shared 4 KB body the loop stays in the instruction cache.
template <class P>
__device__ __noinline__ Mont256<P, false> mont_mul_outlined(Mont256<P, false> a, Mont256<P, false> b);

template <class P, bool COMPACT = false>
struct __align__(16) Mont256 {
  uint32_t v[8];
  using Params = P;
  using Plain = Mont256<P, false>;

  static ZKP_DEVINL Mont256 zero() {
    Mont256 r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = 0;
    return r;
  }
  static ZKP_DEVINL Mont256 one() {
    Mont256 r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = P::R1_(i);
    return r;
  }
  static ZKP_DEVINL Mont256 r2() {
    Mont256 r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = P::R2_(i);
    return r;
  }
  ZKP_DEVINL bool is_zero() const {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= v[i];
    return o == 0;
  }
  ZKP_DEVINL bool operator==(const Mont256& b) const {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= v[i] ^ b.v[i];
    return o == 0;
  }
  ZKP_DEVINL bool operator!=(const Mont256& b) const { return !(*this == b); }

  // r = (t >= MOD) ? t - MOD : t        (t < 2*MOD < 2^256)
  static ZKP_DEVINL void final_sub(uint32_t (&t)[8]) {
    uint32_t s[8], borrow;
    asm("sub.cc.u32  %0, %9,  %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32    %8, 0, 0;"
        : "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(s[3]), "=r"(s[4]), "=r"(s[5]), "=r"(s[6]), "=r"(s[7]),
          "=r"(borrow)
        : "r"(t[0]), "r"(t[1]), "r"(t[2]), "r"(t[3]), "r"(t[4]), "r"(t[5]), "r"(t[6]), "r"(t[7]),
          "r"(P::MOD_(0)), "r"(P::MOD_(1)), "r"(P::MOD_(2)), "r"(P::MOD_(3)), "r"(P::MOD_(4)),
          "r"(P::MOD_(5)), "r"(P::MOD_(6)), "r"(P::MOD_(7)));
#pragma unroll
    for (int i = 0; i < 8; i++) t[i] = borrow ? t[i] : s[i];
  }

  friend ZKP_DEVINL Mont256 operator+(const Mont256& a, const Mont256& b) {
    Mont256 r;
    asm("add.cc.u32  %0, %8,  %16;\n\t"
        "addc.cc.u32 %1, %9,  %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32    %7, %15, %23;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
          "=r"(r.v[7])
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]),
          "r"(a.v[7]), "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]),
          "r"(b.v[6]), "r"(b.v[7]));
    final_sub(r.v);  // a + b < 2*MOD < 2^255: no carry out of limb 7
    return r;
  }

  friend ZKP_DEVINL Mont256 operator-(const Mont256& a, const Mont256& b) {
    Mont256 r;
    uint32_t borrow;
    asm("sub.cc.u32  %0, %9,  %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32    %8, 0, 0;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
          "=r"(r.v[7]), "=r"(borrow)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]),
          "r"(a.v[7]), "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]),
          "r"(b.v[6]), "r"(b.v[7]));
    // borrow is 0 or 0xffffffff: add back MOD & borrow
    asm("add.cc.u32  %0, %0, %8;\n\t"
        "addc.cc.u32 %1, %1, %9;\n\t"
        "addc.cc.u32 %2, %2, %10;\n\t"
        "addc.cc.u32 %3, %3, %11;\n\t"
        "addc.cc.u32 %4, %4, %12;\n\t"
        "addc.cc.u32 %5, %5, %13;\n\t"
        "addc.cc.u32 %6, %6, %14;\n\t"
        "addc.u32    %7, %7, %15;"
        : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]),
          "+r"(r.v[7])
        : "r"(P::MOD_(0) & borrow), "r"(P::MOD_(1) & borrow), "r"(P::MOD_(2) & borrow),
          "r"(P::MOD_(3) & borrow), "r"(P::MOD_(4) & borrow), "r"(P::MOD_(5) & borrow),
          "r"(P::MOD_(6) & borrow), "r"(P::MOD_(7) & borrow));
    return r;
  }

  ZKP_DEVINL Mont256 neg() const { return is_zero() ? *this : (zero() - *this); }
  ZKP_DEVINL Mont256 dbl() const { return *this + *this; }

  friend ZKP_DEVINL Mont256 operator*(const Mont256& a, const Mont256& b) {
    if constexpr (COMPACT) {
      Plain pa, pb;
#pragma unroll
      for (int i = 0; i < 8; i++) {
        pa.v[i] = a.v[i];
        pb.v[i] = b.v[i];
      }
      Plain pr = mont_mul_outlined<P>(pa, pb);
      Mont256 r;
#pragma unroll
      for (int i = 0; i < 8; i++) r.v[i] = pr.v[i];
      return r;
    } else {
      return mul_inline(a, b);
    }
  }

  static ZKP_DEVINL Mont256 mul_inline(const Mont256& a, const Mont256& b) {
    uint32_t E[8], O[8];
#pragma unroll
    for (int i = 0; i < 8; i++) E[i] = O[i] = 0;
    uint32_t left = 0;
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      detail::mont_row<P>(E, O, left, a.v, b.v[i]);
      // E[0] == 0 now; T >>= 32: O becomes the aligned accumulator, E shifts down two limbs.
      left = E[1];
#pragma unroll
      for (int k = 0; k < 6; k++) E[k] = E[k + 2];
      E[6] = E[7] = 0;
      detail::mont_row<P>(O, E, left, a.v, b.v[i + 1]);
      left = O[1];
#pragma unroll
      for (int k = 0; k < 6; k++) O[k] = O[k + 2];
      O[6] = O[7] = 0;
    }
    // T = E + left + (O << 32), O[6] = O[7] = 0, T < 2*MOD
    Mont256 r;
    asm("add.cc.u32  %0, %8,  %16;\n\t"
        "addc.cc.u32 %1, %9,  %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32    %7, %15, %23;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
          "=r"(r.v[7])
        : "r"(E[0]), "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]), "r"(left),
          "r"(O[0]), "r"(O[1]), "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]));
    final_sub(r.v);
    return r;
  }

  // ---- lazy reduction (Fp2 products, fp2.cuh): the 512-bit product and the Montgomery reduction as two steps, so
  // that sums and differences of products are reduced ONCE.  mul_wide is the CIOS loop above without its
  // reduction half (64 limb-MACs; the low limb of every row is final), for any a, b < 2^256.  redc_wide is the
  // reduction half alone (72 limb-MACs): T < MOD * 2^256  ->  T / 2^256 mod MOD, fully reduced.  The upper limbs of
  // T enter the sliding window one per row, in the slot the shift has just vacated.
  static ZKP_DEVINL void mul_wide(const Mont256& a, const Mont256& b, uint32_t (&T)[16]) {
    uint32_t E[8], O[8];
#pragma unroll
    for (int i = 0; i < 8; i++) E[i] = O[i] = 0;
    uint32_t left = 0;
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      detail::mad_lanes_cin(E[0], left, O, a.v[1], a.v[3], a.v[5], a.v[7], b.v[i]);
      detail::mad_lanes_top(E, a.v[0], a.v[2], a.v[4], a.v[6], b.v[i], O[7]);
      T[i] = E[0];
      left = E[1];
#pragma unroll
      for (int k = 0; k < 6; k++) E[k] = E[k + 2];
      E[6] = E[7] = 0;
      detail::mad_lanes_cin(O[0], left, E, a.v[1], a.v[3], a.v[5], a.v[7], b.v[i + 1]);
      detail::mad_lanes_top(O, a.v[0], a.v[2], a.v[4], a.v[6], b.v[i + 1], E[7]);
      T[i + 1] = O[0];
      left = O[1];
#pragma unroll
      for (int k = 0; k < 6; k++) O[k] = O[k + 2];
      O[6] = O[7] = 0;
    }
    // upper half = E + left + (O << 32)
    asm("add.cc.u32  %0, %8,  %16;\n\t"
        "addc.cc.u32 %1, %9,  %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32    %7, %15, %23;"
        : "=r"(T[8]), "=r"(T[9]), "=r"(T[10]), "=r"(T[11]), "=r"(T[12]), "=r"(T[13]), "=r"(T[14]), "=r"(T[15])
        : "r"(E[0]), "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]), "r"(left),
          "r"(O[0]), "r"(O[1]), "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]));
  }

  template <int I>
  static ZKP_DEVINL void redc_row(uint32_t (&X)[8], uint32_t (&Y)[8], uint32_t& left, const uint32_t (&T)[16]) {
    detail::add_cin(X[0], left, Y);
    uint32_t q = X[0] * P::N0INV;
    detail::mad_lanes_top(X, P::MOD_(0), P::MOD_(2), P::MOD_(4), P::MOD_(6), q, Y[7]);
    uint32_t cy = detail::mad_lanes(Y, P::MOD_(1), P::MOD_(3), P::MOD_(5), P::MOD_(7), q);
    (void)cy;
    left = X[1];
#pragma unroll
    for (int k = 0; k < 6; k++) X[k] = X[k + 2];
    X[6] = T[8 + I];  // limb 8 + I of T joins the window in the slot the shift vacated
    X[7] = 0;
  }
  static ZKP_DEVINL Mont256 redc_wide(const uint32_t (&T)[16]) {
    uint32_t E[8], O[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
      E[i] = T[i];
      O[i] = 0;
    }
    uint32_t left = 0;
    redc_row<0>(E, O, left, T);
    redc_row<1>(O, E, left, T);
    redc_row<2>(E, O, left, T);
    redc_row<3>(O, E, left, T);
    redc_row<4>(E, O, left, T);
    redc_row<5>(O, E, left, T);
    redc_row<6>(E, O, left, T);
    redc_row<7>(O, E, left, T);
    Mont256 r;
    asm("add.cc.u32  %0, %8,  %16;\n\t"
        "addc.cc.u32 %1, %9,  %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32    %7, %15, %23;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
          "=r"(r.v[7])
        : "r"(E[0]), "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]), "r"(left),
          "r"(O[0]), "r"(O[1]), "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]));
    final_sub(r.v);
    return r;
  }

  // Dedicated squaring: 36 instead of 64 limb products for a * a (each cross product once, against the
  // doubled operand), same 72 for the reduction: 108 + 8 wide MACs instead of 136.  Intermediate T stays
  // below 3 MOD < 2^256 (a row adds at most a_i * 2a), the final value (a^2 + Q MOD) / 2^256 below 1.25 MOD.
  template <int I>
  static ZKP_DEVINL void sqr_row(uint32_t (&X)[8], uint32_t (&Y)[8], uint32_t& left, const uint32_t (&a)[8],
                                 const uint32_t (&d)[8]) {
    uint32_t m[8];
#pragma unroll
    for (int j = 0; j < 8; j++) m[j] = j == I ? a[j] : (j == I + 1 ? (a[j] << 1) : d[j]);
    detail::mont_row_sq<P, I>(X, Y, left, m, a[I]);
    left = X[1];
#pragma unroll
    for (int k = 0; k < 6; k++) X[k] = X[k + 2];
    X[6] = X[7] = 0;
  }
  static ZKP_DEVINL Mont256 sqr_inline(const Mont256& a) {
    uint32_t d[8];  // limbs of 2a (a < 2^254: nothing is shifted out)
    d[0] = a.v[0] << 1;
#pragma unroll
    for (int j = 1; j < 8; j++) d[j] = (a.v[j] << 1) | (a.v[j - 1] >> 31);
    uint32_t E[8], O[8];
#pragma unroll
    for (int i = 0; i < 8; i++) E[i] = O[i] = 0;
    uint32_t left = 0;
    sqr_row<0>(E, O, left, a.v, d);
    sqr_row<1>(O, E, left, a.v, d);
    sqr_row<2>(E, O, left, a.v, d);
    sqr_row<3>(O, E, left, a.v, d);
    sqr_row<4>(E, O, left, a.v, d);
    sqr_row<5>(O, E, left, a.v, d);
    sqr_row<6>(E, O, left, a.v, d);
    sqr_row<7>(O, E, left, a.v, d);
    Mont256 r;
    asm("add.cc.u32  %0, %8,  %16;\n\t"
        "addc.cc.u32 %1, %9,  %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32    %7, %15, %23;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
          "=r"(r.v[7])
        : "r"(E[0]), "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]), "r"(left),
          "r"(O[0]), "r"(O[1]), "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]));
    final_sub(r.v);
    return r;
  }

  ZKP_DEVINL Mont256 sqr() const {
    if constexpr (COMPACT) return (*this) * (*this);  // small-code flavour: one shared product body
    else return sqr_inline(*this);
  }

  // canonical integer (non-Montgomery) <-> Montgomery
  ZKP_DEVINL Mont256 to_mont() const { return (*this) * r2(); }
  ZKP_DEVINL Mont256 from_mont() const {
    Mont256 o;
#pragma unroll
    for (int i = 0; i < 8; i++) o.v[i] = i == 0 ? 1u : 0u;
    return (*this) * o;
  }

  // a^e for a canonical (non-Montgomery) little-endian exponent held in 8 limbs
  ZKP_DEVINL Mont256 pow(const uint32_t (&e)[8]) const {
    Mont256 acc = one();
    for (int i = 7; i >= 0; i--) {
      for (int b = 31; b >= 0; b--) {
        acc = acc.sqr();
        if ((e[i] >> b) & 1) acc = acc * (*this);
      }
    }
    return acc;
  }

  // Inverse, inv(0) = 0 as in py_ecc's prime_field_inv.  Batched division steps on the integer held in the
  // limbs (modinv30.cuh: 30 steps on 32-bit registers per 2x2 matrix applied to the 270-bit values), then one
  // Montgomery product by R^3 to land back in Montgomery form: (aR)^-1 * R^3 * R^-1 = a^-1 R.
  __device__ __noinline__ Mont256 inv() const {
    if (is_zero()) return *this;
    Mont256 r, r3;
    modinv30::inverse<P>(v, r.v);
#pragma unroll
    for (int i = 0; i < 8; i++) r3.v[i] = P::R3_(i);
    return r * r3;
  }

  // The inversion of rounds 1-2: binary extended Euclid (shifts, adds and subtractions only: ~4x shorter than the
  // 254-squaring Fermat chain for a lone thread, ~2.5x longer than inv()).  Kept as an independent cross-check.
  __device__ __noinline__ Mont256 inv_euclid() const {
    if (is_zero()) return *this;
    uint32_t u[8], w[8], x1[8], x2[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
      u[i] = v[i];
      w[i] = P::MOD_(i);
      x1[i] = i == 0 ? 1u : 0u;
      x2[i] = 0u;
    }
    auto is_one = [](const uint32_t (&a)[8]) {
      uint32_t o = a[0] ^ 1u;
#pragma unroll
      for (int i = 1; i < 8; i++) o |= a[i];
      return o == 0;
    };
    auto shr1 = [](uint32_t (&a)[8], uint32_t top) {
#pragma unroll
      for (int i = 0; i < 7; i++) a[i] = (a[i] >> 1) | (a[i + 1] << 31);
      a[7] = (a[7] >> 1) | (top << 31);
    };
    auto halve_mod = [&](uint32_t (&a)[8]) {  // a <- a / 2 mod MOD
      uint32_t carry = 0;
      if (a[0] & 1u) {
        uint64_t c = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
          c += (uint64_t)a[i] + P::MOD_(i);
          a[i] = (uint32_t)c;
          c >>= 32;
        }
        carry = (uint32_t)c;
      }
      shr1(a, carry);
    };
    auto geq = [](const uint32_t (&a)[8], const uint32_t (&b)[8]) {
#pragma unroll
      for (int i = 7; i >= 0; i--) {
        if (a[i] != b[i]) return a[i] > b[i];
      }
      return true;
    };
    auto sub = [](uint32_t (&a)[8], const uint32_t (&b)[8]) {  // a <- a - b, returns borrow
      uint64_t br = 0;
#pragma unroll
      for (int i = 0; i < 8; i++) {
        uint64_t d = (uint64_t)a[i] - b[i] - br;
        a[i] = (uint32_t)d;
        br = (d >> 63) & 1u;
      }
      return (uint32_t)br;
    };
    auto sub_mod = [&](uint32_t (&a)[8], const uint32_t (&b)[8]) {  // a <- a - b mod MOD
      if (sub(a, b)) {
        uint64_t c = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
          c += (uint64_t)a[i] + P::MOD_(i);
          a[i] = (uint32_t)c;
          c >>= 32;
        }
      }
    };
    while (!is_one(u) && !is_one(w)) {
      while (!(u[0] & 1u)) {
        shr1(u, 0);
        halve_mod(x1);
      }
      while (!(w[0] & 1u)) {
        shr1(w, 0);
        halve_mod(x2);
      }
      if (geq(u, w)) {
        sub(u, w);
        sub_mod(x1, x2);
      } else {
        sub(w, u);
        sub_mod(x2, x1);
      }
    }
    Mont256 r, r3;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      r.v[i] = is_one(u) ? x1[i] : x2[i];
      r3.v[i] = P::R3_(i);
    }
    return r * r3;
  }

  // Fermat inverse a^(MOD-2), kept as an independent cross-check of inv() (tests).
  __device__ __noinline__ Mont256 inv_fermat() const {
    uint32_t e[8];
#pragma unroll
    for (int i = 0; i < 8; i++) e[i] = P::MOD_(i);
    e[0] -= 2;  // MOD is odd and its low limb is >= 2 for both fields
    return pow(e);
  }
};

template <class P>
__device__ __noinline__ Mont256<P, false> mont_mul_outlined(Mont256<P, false> a, Mont256<P, false> b) {
  return Mont256<P, false>::mul_inline(a, b);
}
template <class P>
__device__ __noinline__ Mont256<P, false> mont_sqr_outlined(Mont256<P, false> a) {
  return Mont256<P, false>::sqr_inline(a);
}

using Fp = Mont256<FpParams>;
using Fr = Mont256<FrParams>;
using FpC = Mont256<FpParams, true>;  // same layout as Fp, out-of-line products
using FrC = Mont256<FrParams, true>;

}  // namespace zkp
