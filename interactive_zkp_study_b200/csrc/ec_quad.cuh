// XYZZ group operations spread over four adjacent lanes (a "quad").
//
// The reduction tail of the MSM (bucket fold, weighted-sum levels) is a few thousand group operations
// deep in dependency, not wide: it is bounded by the latency of ONE addition on ONE thread, and a lone
// thread cannot overlap independent Montgomery products (zkp_latency_probe: 437 ns for one product,
// 877 ns for two independent ones -- separate carry chains are not interleaved by ptxas).  An XYZZ
// addition is 14 products in 4 dependent levels (doubling: 9 in 3), so four lanes that each take one
// product per level and exchange the results by warp shuffles finish it in ~4 product latencies
// instead of 14.  Every lane keeps the full operands and the full result (the additions /
// subtractions between the levels are recomputed by all four lanes: they are cheap), which keeps
// the control flow warp-uniform: ALL 32 lanes of a warp must call these functions together.
#pragma once
#include "ec.cuh"

namespace zkp {

template <class F>
struct Quad {
  static constexpr int WORDS = sizeof(F) / 4;
  static constexpr unsigned FULL = 0xffffffffu;

  // value of `x` held by lane `src` (0..3) of the caller's quad
  static ZKP_DEVINL F from(const F& x, int src) {
    F r;
    const uint32_t* in = reinterpret_cast<const uint32_t*>(&x);
    uint32_t* out = reinterpret_cast<uint32_t*>(&r);
    const int lane = (threadIdx.x & 31 & ~3) | src;
#pragma unroll
    for (int i = 0; i < WORDS; i++) out[i] = __shfl_sync(FULL, in[i], lane);
    return r;
  }
  // lane ql of the quad takes a / b / c / d.  Written with masks: as nested conditionals ptxas turned the choice
  // into four divergent paths per call (16 BSSY / BRA / BSYNC per addition, the lanes of a quad executing their
  // copies one after the other) and the glue between the products cost 3.3 us of a 5.2 us addition.
  static ZKP_DEVINL F sel(int ql, const F& a, const F& b, const F& c, const F& d) {
    F r;
    const uint32_t* pa = reinterpret_cast<const uint32_t*>(&a);
    const uint32_t* pb = reinterpret_cast<const uint32_t*>(&b);
    const uint32_t* pc = reinterpret_cast<const uint32_t*>(&c);
    const uint32_t* pd = reinterpret_cast<const uint32_t*>(&d);
    uint32_t* out = reinterpret_cast<uint32_t*>(&r);
    const uint32_t m1 = 0u - ((uint32_t)ql & 1u), m2 = 0u - (((uint32_t)ql >> 1) & 1u);
#pragma unroll
    for (int i = 0; i < WORDS; i++) {
      const uint32_t lo = (pa[i] & ~m1) | (pb[i] & m1), hi = (pc[i] & ~m1) | (pd[i] & m1);
      out[i] = (lo & ~m2) | (hi & m2);
    }
    return r;
  }

  // 2 * a on the quad (EFD dbl-2008-s-1, a = 0); 2 * infinity = infinity falls out (ZZ stays 0).
  static ZKP_DEVINL XYZZ<F> dbl(const XYZZ<F>& a, int ql) {
    F u = a.y.dbl();
    // level 1: V = U^2 | XX = X^2
    F m1 = sel(ql, u, a.x, u, a.x);
    m1 = m1 * m1;
    F v = from(m1, 0), xx = from(m1, 1);
    F m = xx.dbl() + xx;
    // level 2: W = U*V | S = X*V | MM = M^2 | ZZ3 = V*ZZ
    F m2 = sel(ql, u, a.x, m, a.zz) * sel(ql, v, v, m, v);
    F w = from(m2, 0), s = from(m2, 1), mm = from(m2, 2);
    XYZZ<F> r;
    r.x = mm - s.dbl();
    r.zz = from(m2, 3);
    // level 3: M*(S - X3) | W*Y | W*ZZZ
    F m3 = sel(ql, m, w, w, w) * sel(ql, s - r.x, a.y, a.zzz, a.zzz);
    r.y = from(m3, 0) - from(m3, 1);
    r.zzz = from(m3, 2);
    return r;
  }

  // a + b on the quad (EFD add-2008-s) with the identity / doubling / inverse cases of XYZZ::add.
  static ZKP_DEVINL XYZZ<F> add(const XYZZ<F>& a, const XYZZ<F>& b, int ql) {
    // level 1: U1 = X1*ZZ2 | U2 = X2*ZZ1 | S1 = Y1*ZZZ2 | S2 = Y2*ZZZ1
    F m1 = sel(ql, a.x, b.x, a.y, b.y) * sel(ql, b.zz, a.zz, b.zzz, a.zzz);
    F u1 = from(m1, 0), u2 = from(m1, 1), s1 = from(m1, 2), s2 = from(m1, 3);
    F p = u2 - u1, rr = s2 - s1;
    // level 2: PP = P^2 | ZZ12 = ZZ1*ZZ2 | RR = R^2 | ZZZ12 = ZZZ1*ZZZ2
    F m2 = sel(ql, p, a.zz, rr, a.zzz) * sel(ql, p, b.zz, rr, b.zzz);
    F pp = from(m2, 0), r2 = from(m2, 2);
    // level 3: PPP = P*PP | ZZ3 = ZZ12*PP | Q = U1*PP | (idle)
    F m3 = sel(ql, p, m2, u1, p) * pp;
    F ppp = from(m3, 0), q = from(m3, 2);
    XYZZ<F> r;
    r.x = r2 - ppp - q.dbl();
    r.zz = from(m3, 1);
    // level 4: S1*PPP | (idle) | R*(Q - X3) | ZZZ3 = ZZZ12*PPP
    F m4 = sel(ql, s1, s1, rr, m2) * sel(ql, ppp, ppp, q - r.x, ppp);
    r.y = from(m4, 2) - from(m4, 0);
    r.zzz = from(m4, 3);
    // edge cases, decided identically by the four lanes; the doubling fallback is taken by the whole
    // warp together (it shuffles), and only when some quad needs it
    const bool a_inf = a.is_inf(), b_inf = b.is_inf();
    const bool same_x = !a_inf && !b_inf && p.is_zero();
    const bool need_dbl = same_x && rr.is_zero();
    if (__any_sync(FULL, need_dbl)) {
      XYZZ<F> d = dbl(a, ql);
      if (need_dbl) r = d;
    }
    if (same_x && !need_dbl) r = XYZZ<F>::inf();
    if (b_inf) r = a;
    if (a_inf) r = b;
    return r;
  }
};

}  // namespace zkp
