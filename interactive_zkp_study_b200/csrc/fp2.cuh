// Fp2 = Fp[u]/(u^2 + 1) for the BN254 G2 twist.  Device-side replacement for py_ecc's FQ2
// (call sites: /root/reference/zkp/groth16/proving.py:35-45 proof_b; setup.py:62-69;
// wire layout /root/reference/plonk_serializers.py:56-67 -> coeffs[0] + coeffs[1]*u).
#pragma once
#include "ff.cuh"

namespace zkp {

struct __align__(16) Fp2 {
  // Components use the out-of-line Fp product (arguments and result in registers): an inlined Fp2
  // product is three calls plus a few adds, so the G2 kernels stay small (fast to compile, no stack
  // traffic for operands) while the multiplier body is shared.
  using Fp = FpC;
  Fp c0, c1;

  static ZKP_DEVINL Fp2 zero() { return {Fp::zero(), Fp::zero()}; }
  static ZKP_DEVINL Fp2 one() { return {Fp::one(), Fp::zero()}; }
  ZKP_DEVINL bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
  ZKP_DEVINL bool operator==(const Fp2& b) const { return c0 == b.c0 && c1 == b.c1; }
  ZKP_DEVINL bool operator!=(const Fp2& b) const { return !(*this == b); }
  friend ZKP_DEVINL Fp2 operator+(const Fp2& a, const Fp2& b) { return {a.c0 + b.c0, a.c1 + b.c1}; }
  friend ZKP_DEVINL Fp2 operator-(const Fp2& a, const Fp2& b) { return {a.c0 - b.c0, a.c1 - b.c1}; }
  ZKP_DEVINL Fp2 neg() const { return {c0.neg(), c1.neg()}; }
  ZKP_DEVINL Fp2 dbl() const { return {c0.dbl(), c1.dbl()}; }
  // Karatsuba: 3 Fp multiplications
  friend ZKP_DEVINL Fp2 operator*(const Fp2& a, const Fp2& b) {
    Fp t0 = a.c0 * b.c0;
    Fp t1 = a.c1 * b.c1;
    Fp t2 = (a.c0 + a.c1) * (b.c0 + b.c1);
    return {t0 - t1, t2 - t0 - t1};
  }
  // (c0 + c1 u)^2 = (c0 + c1)(c0 - c1) + 2 c0 c1 u : 2 Fp multiplications
  ZKP_DEVINL Fp2 sqr() const {
    Fp s = (c0 + c1) * (c0 - c1);
    Fp m = c0 * c1;
    return {s, m.dbl()};
  }
  // 1 / (c0 + c1 u) = (c0 - c1 u) / (c0^2 + c1^2); inv(0) = 0
  __device__ __noinline__ Fp2 inv() const {
    Fp d = (c0.sqr() + c1.sqr()).inv();
    return {c0 * d, (c1 * d).neg()};
  }
  ZKP_DEVINL Fp2 to_mont() const { return {c0.to_mont(), c1.to_mont()}; }
  ZKP_DEVINL Fp2 from_mont() const { return {c0.from_mont(), c1.from_mont()}; }
};

}  // namespace zkp
