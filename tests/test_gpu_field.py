"""-m gpu: the Montgomery field core and the XYZZ group law of libzkp_b200 against Python ints
(field ops) and the plain-int oracle (group law, incl. every edge case py_ecc defines:
infinity operands, P+P, P+(-P))."""
import random

import pytest

from oracle import bn254

pytestmark = pytest.mark.gpu

P, R = bn254.P, bn254.R


def _vec(vals):
    return b"".join(int(v).to_bytes(32, "little") for v in vals)


def _unvec(b):
    return [int.from_bytes(b[i:i + 32], "little") for i in range(0, len(b), 32)]


@pytest.mark.parametrize("field,mod", [(0, P), (1, R)])
def test_field_ops_bit_exact(native, field, mod):
    rng = random.Random(1234 + field)
    edge = [0, 1, 2, mod - 1, mod - 2, (1 << 253), (1 << 32) - 1, 1 << 32, (mod - 1) // 2, (mod + 1) // 2]
    a = edge + [rng.randrange(mod) for _ in range(4000)]
    b = list(reversed(edge)) + [rng.randrange(mod) for _ in range(4000)]
    n = len(a)
    ab, bb = _vec(a), _vec(b)
    assert _unvec(native.dbg_field_op(field, 0, ab, bb, n)) == [(x + y) % mod for x, y in zip(a, b)]
    assert _unvec(native.dbg_field_op(field, 1, ab, bb, n)) == [(x - y) % mod for x, y in zip(a, b)]
    assert _unvec(native.dbg_field_op(field, 2, ab, bb, n)) == [(x * y) % mod for x, y in zip(a, b)]
    assert _unvec(native.dbg_field_op(field, 4, ab, None, n)) == [(x * x) % mod for x in a]
    # limb patterns IN MONTGOMERY FORM (the device squares x = a R mod p): all-ones low limbs, top bits of
    # every limb set (the dedicated squaring doubles the operand across limb boundaries), single-limb values
    rinv = pow(1 << 256, -1, mod)
    top = mod >> 224
    pats = [mod - 1, (1 << 224) - 1, ((top - 1) << 224) | ((1 << 224) - 1), int("80000000" * 7, 16),
            (0x20000000 << 224) | int("80000000" * 7, 16), int("ffffffff" * 7, 16) ^ int("0000ffff" * 7, 16),
            0xffffffff, 0xffffffff << 32, 0xffffffff << 192, 0x80000000, (0x80000000 << 96) | 0x80000000,
            (1 << 253) + (1 << 31), int("7fffffff" * 7, 16), int("0000000180000000" * 3, 16)]
    pats = [x % mod for x in pats]
    pa = [x * rinv % mod for x in pats]
    assert _unvec(native.dbg_field_op(field, 4, _vec(pa), None, len(pa))) == [(x * x) % mod for x in pa]
    assert _unvec(native.dbg_field_op(field, 2, _vec(pa), _vec(pa[::-1]), len(pa))) == [(x * y) % mod for x, y in zip(pa, pa[::-1])]
    # the two-step form (512-bit product, then one Montgomery reduction) behind the lazily reduced Fp2 product
    assert _unvec(native.dbg_field_op(field, 6, _vec(pa), _vec(pa[::-1]), len(pa))) == [(x * y) % mod for x, y in zip(pa, pa[::-1])]
    assert _unvec(native.dbg_field_op(field, 6, _vec(a[:500]), _vec(b[:500]), 500)) == [(x * y) % mod for x, y in zip(a[:500], b[:500])]
    # a b - c d with one reduction (the Y coordinate of every XYZZ formula): c = a + b, d = a - b, random and limb patterns
    want = lambda xs, ys: [(x * y - (x + y) * (x - y)) % mod for x, y in zip(xs, ys)]
    assert _unvec(native.dbg_field_op(field, 8, ab, bb, n)) == want(a, b)
    assert _unvec(native.dbg_field_op(field, 8, _vec(pa), _vec(pa[::-1]), len(pa))) == want(pa, pa[::-1])
    assert _unvec(native.dbg_field_op(field, 8, _vec(pats), _vec(pats[::-1]), len(pats))) == want(pats, pats[::-1])
    m = 300
    got = _unvec(native.dbg_field_op(field, 3, _vec(a[:m]), None, m))
    assert got == [bn254.inv(x, mod) for x in a[:m]]  # inv(0) == 0 as in py_ecc (batched division steps, modinv30.cuh)
    assert _unvec(native.dbg_field_op(field, 5, _vec(a[:m]), None, m)) == got  # Fermat chain agrees
    assert _unvec(native.dbg_field_op(field, 7, _vec(a[:m]), None, m)) == got  # binary extended Euclid agrees
    # values whose Montgomery form a R mod p is small / has long runs of zeros or ones (the division steps look at
    # the low bits of that integer), and small plain values
    mont_small = [x * rinv % mod for x in (1, 2, 3, 1 << 30, (1 << 30) - 1, 1 << 60, (1 << 255) % mod, mod - 1, mod - 2,
                                            (1 << 200) - 1, int("aaaaaaaa" * 7, 16))]
    more = mont_small + [1, 2, 3, 5, 1 << 30, 1 << 31, 1 << 32, mod - 1, mod - 2, (mod + 1) // 2] + [rng.randrange(1 << k) for k in range(1, 254, 3)]
    got = _unvec(native.dbg_field_op(field, 3, _vec(more), None, len(more)))
    assert got == [bn254.inv(x, mod) for x in more]


def test_fp2_ops_bit_exact(native):
    """Fp2 product (Karatsuba on unreduced 512-bit products, one reduction per component), squaring, inverse."""
    rng = random.Random(99)
    edge = [0, 1, 2, P - 1, P - 2, (P - 1) // 2, (P + 1) // 2, (1 << 253), (1 << 32) - 1]
    xs = [(a, b) for a in edge for b in edge] + [(rng.randrange(P), rng.randrange(P)) for _ in range(3000)]
    ys = [(b, a) for a, b in reversed(xs[:81])] + [(rng.randrange(P), rng.randrange(P)) for _ in range(3000)]
    enc = lambda v: b"".join(int(c).to_bytes(32, "little") for pair in v for c in pair)
    dec = lambda b: [(int.from_bytes(b[64 * i:64 * i + 32], "little"), int.from_bytes(b[64 * i + 32:64 * i + 64], "little"))
                     for i in range(len(b) // 64)]
    n = len(xs)
    assert dec(native.dbg_field_op(2, 2, enc(xs), enc(ys), n)) == [bn254.f2_mul(x, y) for x, y in zip(xs, ys)]
    assert dec(native.dbg_field_op(2, 4, enc(xs), None, n)) == [bn254.f2_mul(x, x) for x in xs]
    inv = dec(native.dbg_field_op(2, 3, enc(xs[81:381]), None, 300))
    assert all(bn254.f2_mul(x, y) == (1, 0) for x, y in zip(xs[81:381], inv))


@pytest.mark.parametrize("quad", [0, 2])
def test_g1_add_all_cases(native, quad):
    """quad = 2: the same sums through the four-lane group operations of ec_quad.cuh (add, double, and the
    P + P / P + (-P) / infinity cases inside them)."""
    rng = random.Random(7)
    pts = [bn254.g1_mul(bn254.G1, rng.randrange(1, R)) for _ in range(40)]
    a, b = [], []
    for i in range(0, 40, 2):
        a.append(pts[i]); b.append(pts[i + 1])          # generic
    a += [None, pts[0], None, pts[1], pts[2]]
    b += [pts[0], None, None, pts[1], bn254.g1_neg(pts[2])]  # inf+P, P+inf, inf+inf, P+P, P+(-P)
    n = len(a)
    out = native.dbg_point_add(quad, native.g1_vec_bytes(a), native.g1_vec_bytes(b), n)
    got = [native.g1_from_bytes(out[64 * i:64 * i + 64]) for i in range(n)]
    assert got == [bn254.g1_add(x, y) for x, y in zip(a, b)]


@pytest.mark.parametrize("quad", [1, 3])
def test_g2_add_all_cases(native, quad):
    rng = random.Random(8)
    pts = [bn254.g2_mul(bn254.G2, rng.randrange(1, 1 << 64)) for _ in range(12)]
    a = pts[0:6] + [None, pts[0], None, pts[1], pts[2]]
    b = pts[6:12] + [pts[0], None, None, pts[1], bn254.g2_neg(pts[2])]
    n = len(a)
    enc = lambda v: b"".join(native.g2_bytes(p) for p in v)
    out = native.dbg_point_add(quad, enc(a), enc(b), n)
    got = [native.g2_from_bytes(out[128 * i:128 * i + 128]) for i in range(n)]
    assert got == [bn254.g2_add(x, y) for x, y in zip(a, b)]
