// BN254 group law on the device: y^2 = x^3 + 3 over Fp (G1) and y^2 = x^3 + 3/(9+u) over
// Fp2 (G2).  Replaces py_ecc's affine bn128.add / double / multiply / neg as called from
// /root/reference/zkp/groth16/proving.py:12-15,27-31 and /root/reference/zkp/plonk/field.py:88,103.
//
// py_ecc performs one field inversion per affine add.  Here bucket sums are carried in XYZZ
// coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2; EFD "xyzz", a = 0): mixed add 8M+2S, full add
// 12M+2S, doubling 6M+4S, and a single inversion per MSM converts the result back to the unique
// affine representative, so outputs are bit-identical with the reference's.
//
// Edge semantics restated from py_ecc: infinity is an absorbing-free identity
// (add(None,P)=P), P+P doubles, P+(-P)=infinity.  Affine infinity is encoded as (0,0), which is
// not on either curve; XYZZ infinity is ZZ == 0.
#pragma once
#include "fp2.cuh"

namespace zkp {

template <class F>
struct __align__(16) Affine {
  F x, y;
  ZKP_DEVINL bool is_inf() const { return x.is_zero() && y.is_zero(); }
  static ZKP_DEVINL Affine inf() { return {F::zero(), F::zero()}; }
  ZKP_DEVINL Affine neg() const { return {x, y.neg()}; }
};

// a b - c d: the Y coordinate of every XYZZ formula.  The compact Fp flavour (reduction tail of the MSM, tables)
// takes both products through ONE Montgomery reduction (fp2.cuh fp_mulsub_outlined: 200 instead of 272 limb-MACs).
template <class F>
ZKP_DEVINL F field_mulsub(const F& a, const F& b, const F& c, const F& d) {
  return a * b - c * d;
}
template <>
ZKP_DEVINL FpC field_mulsub<FpC>(const FpC& a, const FpC& b, const FpC& c, const FpC& d) {
  Fp pa, pb, pc, pd;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    pa.v[i] = a.v[i];
    pb.v[i] = b.v[i];
    pc.v[i] = c.v[i];
    pd.v[i] = d.v[i];
  }
  Fp pr = fp_mulsub_outlined(pa, pb, pc, pd);
  FpC r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = pr.v[i];
  return r;
}

template <class F>
struct __align__(16) XYZZ {
  F x, y, zz, zzz;

  static ZKP_DEVINL XYZZ inf() { return {F::zero(), F::zero(), F::zero(), F::zero()}; }
  ZKP_DEVINL bool is_inf() const { return zz.is_zero(); }
  static ZKP_DEVINL XYZZ from_affine(const Affine<F>& p) {
    if (p.is_inf()) return inf();
    return {p.x, p.y, F::one(), F::one()};
  }
  ZKP_DEVINL XYZZ neg() const { return {x, y.neg(), zz, zzz}; }

  // 2 * (affine p), p != infinity.  EFD mdbl-2008-s-1 with a = 0.
  static ZKP_DEVINL XYZZ dbl_affine(const Affine<F>& p) {
    F u = p.y.dbl();
    F v = u.sqr();
    F w = u * v;
    F s = p.x * v;
    F xx = p.x.sqr();
    F m = xx.dbl() + xx;
    XYZZ r;
    r.x = m.sqr() - s.dbl();
    r.y = field_mulsub(m, s - r.x, w, p.y);
    r.zz = v;
    r.zzz = w;
    return r;
  }

  // EFD dbl-2008-s-1, a = 0.  2*infinity = infinity falls out (ZZ stays 0); a point with y == 0
  // cannot exist on these prime-order curves.
  ZKP_DEVINL XYZZ dbl() const {
    F u = y.dbl();
    F v = u.sqr();
    F w = u * v;
    F s = x * v;
    F xx = x.sqr();
    F m = xx.dbl() + xx;
    XYZZ r;
    r.x = m.sqr() - s.dbl();
    r.y = field_mulsub(m, s - r.x, w, y);
    r.zz = v * zz;
    r.zzz = w * zzz;
    return r;
  }

  // this += affine p.  EFD madd-2008-s (8M + 2S) plus the identity / doubling / inverse cases.
  ZKP_DEVINL void madd(const Affine<F>& p) {
    if (p.is_inf()) return;
    if (is_inf()) {
      x = p.x; y = p.y; zz = F::one(); zzz = F::one();
      return;
    }
    F u2 = p.x * zz;
    F s2 = p.y * zzz;
    F pp_ = u2 - x;
    F r = s2 - y;
    if (pp_.is_zero()) {
      if (r.is_zero()) *this = dbl_affine(p);
      else *this = inf();
      return;
    }
    F pp = pp_.sqr();
    F ppp = pp_ * pp;
    F q = x * pp;
    F x3 = r.sqr() - ppp - q.dbl();
    y = field_mulsub(r, q - x3, y, ppp);
    x = x3;
    zz = zz * pp;
    zzz = zzz * ppp;
  }

  // this += o.  EFD add-2008-s (12M + 2S) plus edge cases.
  ZKP_DEVINL void add(const XYZZ& o) {
    if (o.is_inf()) return;
    if (is_inf()) { *this = o; return; }
    F u1 = x * o.zz;
    F u2 = o.x * zz;
    F s1 = y * o.zzz;
    F s2 = o.y * zzz;
    F pp_ = u2 - u1;
    F r = s2 - s1;
    if (pp_.is_zero()) {
      if (r.is_zero()) *this = dbl();
      else *this = inf();
      return;
    }
    F pp = pp_.sqr();
    F ppp = pp_ * pp;
    F q = u1 * pp;
    F x3 = r.sqr() - ppp - q.dbl();
    y = field_mulsub(r, q - x3, s1, ppp);
    x = x3;
    zz = zz * o.zz * pp;
    zzz = zzz * o.zzz * ppp;
  }

  // Unique affine representative (Montgomery form); infinity -> (0,0).
  ZKP_DEVINL Affine<F> to_affine() const {
    if (is_inf()) return Affine<F>::inf();
    F i = (zz * zzz).inv();
    F izz = i * zzz;   // 1/ZZ
    F izzz = i * zz;   // 1/ZZZ
    return {x * izz, y * izzz};
  }
};

// Field flavour with out-of-line products and the same memory layout (see ff.cuh, COMPACT).
template <class F>
struct CompactOf {
  using type = F;
};
template <>
struct CompactOf<Fp> {
  using type = FpC;
};

using G1Affine = Affine<Fp>;
using G2Affine = Affine<Fp2>;
using G1XYZZ = XYZZ<Fp>;
using G2XYZZ = XYZZ<Fp2>;

}  // namespace zkp
