"""TEST INFRASTRUCTURE (oracle) -- plain-int restatement of the BN254 group law.

Same algorithm as oracle/shim/py_ecc/bn128 (affine double-and-add, one modular inversion per
step; py_ecc call sites /root/reference/zkp/groth16/proving.py:12-15,
/root/reference/zkp/plonk/field.py:88,103) but on Python ints / int pairs instead of FQ / FQ2
objects, so that parity checks at 2^10..2^14 points finish in seconds.  Pinned against the shim
(which is pinned by the reference's own 469 tests) in tests/test_oracle.py.

G1 point: (x, y) ints or None.  G2 point: ((x0, x1), (y0, y1)) or None, element a0 + a1*u, u^2 = -1.
"""

P = 21888242871839275222246405745257275088696311157297823662689037894645226208583
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
G1 = (1, 2)
G2 = (
    (10857046999023057135944570762232829481370756359578518086990519993285655852781,
     11559732032986387107991004021392285783925812861821192530917403151452391805634),
    (8495653923123431417604973247489272438418190587263600148770280649306958101930,
     4082367875863433681332203403145435568316851327593401208105741076214120093531),
)


def inv(a, m=P):
    """py_ecc prime_field_inv semantics: inv(0) = 0."""
    a %= m
    return pow(a, -1, m) if a else 0


# ---------------------------------------------------------------- G1
def g1_is_on_curve(pt):
    if pt is None:
        return True
    x, y = pt
    return (y * y - x * x * x - 3) % P == 0


def g1_double(pt):
    if pt is None:
        return None
    x, y = pt
    m = 3 * x * x * inv(2 * y) % P
    nx = (m * m - 2 * x) % P
    ny = (-m * nx + m * x - y) % P
    return (nx, ny)


def g1_add(p1, p2):
    if p1 is None or p2 is None:
        return p1 if p2 is None else p2
    x1, y1 = p1
    x2, y2 = p2
    if x2 == x1 and y2 == y1:
        return g1_double(p1)
    if x2 == x1:
        return None
    m = (y2 - y1) * inv(x2 - x1) % P
    nx = (m * m - x1 - x2) % P
    ny = (-m * nx + m * x1 - y1) % P
    return (nx, ny)


def g1_neg(pt):
    if pt is None:
        return None
    return (pt[0], (-pt[1]) % P)


def g1_mul(pt, n):
    """bn128.multiply: MSB-first double-and-add, no reduction of n (iterative form of the recursion)."""
    if n == 0 or pt is None:
        return None
    acc = None
    for bit in bin(n)[2:]:
        acc = g1_double(acc)
        if bit == "1":
            acc = g1_add(acc, pt)
    return acc


def g1_msm(points, scalars):
    """sum_i scalars[i]*points[i] exactly as kzg.commit does it (kzg.py:59-67): skip zeros, reduce mod r."""
    acc = None
    for pt, s in zip(points, scalars):
        s = int(s) % R
        if s == 0:
            continue
        acc = g1_add(acc, g1_mul(pt, s))
    return acc


# ---------------------------------------------------------------- Fp2 / G2
def f2_add(a, b):
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def f2_sub(a, b):
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


def f2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def f2_scalar(a, k):
    return (a[0] * k % P, a[1] * k % P)


def f2_inv(a):
    d = inv(a[0] * a[0] + a[1] * a[1])
    return (a[0] * d % P, -a[1] * d % P)


B2 = f2_mul((3, 0), f2_inv((9, 1)))


def g2_is_on_curve(pt):
    if pt is None:
        return True
    x, y = pt
    return f2_sub(f2_mul(y, y), f2_mul(f2_mul(x, x), x)) == B2


def g2_double(pt):
    if pt is None:
        return None
    x, y = pt
    m = f2_mul(f2_scalar(f2_mul(x, x), 3), f2_inv(f2_scalar(y, 2)))
    nx = f2_sub(f2_mul(m, m), f2_scalar(x, 2))
    ny = f2_sub(f2_mul(m, f2_sub(x, nx)), y)
    return (nx, ny)


def g2_add(p1, p2):
    if p1 is None or p2 is None:
        return p1 if p2 is None else p2
    x1, y1 = p1
    x2, y2 = p2
    if x2 == x1 and y2 == y1:
        return g2_double(p1)
    if x2 == x1:
        return None
    m = f2_mul(f2_sub(y2, y1), f2_inv(f2_sub(x2, x1)))
    nx = f2_sub(f2_sub(f2_mul(m, m), x1), x2)
    ny = f2_sub(f2_mul(m, f2_sub(x1, nx)), y1)
    return (nx, ny)


def g2_neg(pt):
    if pt is None:
        return None
    return (pt[0], ((-pt[1][0]) % P, (-pt[1][1]) % P))


def g2_mul(pt, n):
    if n == 0 or pt is None:
        return None
    acc = None
    for bit in bin(n)[2:]:
        acc = g2_double(acc)
        if bit == "1":
            acc = g2_add(acc, pt)
    return acc


def g2_msm(points, scalars):
    acc = None
    for pt, s in zip(points, scalars):
        s = int(s) % R
        if s == 0:
            continue
        acc = g2_add(acc, g2_mul(pt, s))
    return acc
