"""TEST INFRASTRUCTURE (oracle) -- restatement of ``py_ecc.bn128`` (py-ecc==7.0.1).

Third-party boundary of the reference's hot path.  Every
group operation the reference performs goes through this surface:
/root/reference/zkp/groth16/proving.py:12-18, setup.py:7-10, verifying.py:11-14,
/root/reference/zkp/plonk/field.py:63-66,88,103,115,138.

Restated semantics of the non-optimised bn128 module: affine short-Weierstrass
points as 2-tuples, the point at infinity is ``None``; ``add(None, P) = P``,
``add(P, P) = double(P)``, ``add(P, -P) = None``; ``multiply(P, 0) = None``,
``multiply(P, 1) = P`` and otherwise MSB-recursive double-and-add with no
reduction of the scalar; ``pairing(Q in G2, P in G1)`` is the optimal-Ate Miller
loop followed by the final exponentiation ``(p^12 - 1) / r`` in FQ12.
"""

from ..fields import (  # noqa: F401
    bn128_FQ as FQ,
    bn128_FQ2 as FQ2,
    bn128_FQ12 as FQ12,
    bn128_FQP as FQP,
    field_modulus,
)

curve_order = 21888242871839275222246405745257275088548364400416034343698204186575808495617

# Curve is y**2 = x**3 + 3
b = FQ(3)
# Twisted curve over FQ**2
b2 = FQ2([3, 0]) / FQ2([9, 1])
# Extension curve over FQ**12; same b value as over FQ
b12 = FQ12([3] + [0] * 11)

# Generator for curve over FQ
G1 = (FQ(1), FQ(2))
# Generator for twisted curve over FQ2
G2 = (
    FQ2(
        [
            10857046999023057135944570762232829481370756359578518086990519993285655852781,
            11559732032986387107991004021392285783925812861821192530917403151452391805634,
        ]
    ),
    FQ2(
        [
            8495653923123431417604973247489272438418190587263600148770280649306958101930,
            4082367875863433681332203403145435568316851327593401208105741076214120093531,
        ]
    ),
)
# Point at infinity over FQ
Z1 = None
# Point at infinity for twisted curve over FQ2
Z2 = None


def is_inf(pt):
    return pt is None


def is_on_curve(pt, b):
    if is_inf(pt):
        return True
    x, y = pt
    return y ** 2 - x ** 3 == b


def double(pt):
    if is_inf(pt):
        return pt
    x, y = pt
    m = 3 * x ** 2 / (2 * y)
    newx = m ** 2 - 2 * x
    newy = -m * newx + m * x - y
    return (newx, newy)


def add(p1, p2):
    if p1 is None or p2 is None:
        return p1 if p2 is None else p2
    x1, y1 = p1
    x2, y2 = p2
    if x2 == x1 and y2 == y1:
        return double(p1)
    elif x2 == x1:
        return None
    else:
        m = (y2 - y1) / (x2 - x1)
    newx = m ** 2 - x1 - x2
    newy = -m * newx + m * x1 - y1
    assert newy == (-m * newx + m * x2 - y2)
    return (newx, newy)


def multiply(pt, n):
    if n == 0:
        return None
    elif n == 1:
        return pt
    elif not n % 2:
        return multiply(double(pt), n // 2)
    else:
        return add(multiply(double(pt), int(n // 2)), pt)


def eq(p1, p2):
    return p1 == p2


# "Twist" a point in E(FQ2) into a point in E(FQ12)
w = FQ12([0, 1] + [0] * 10)


def neg(pt):
    if pt is None:
        return None
    x, y = pt
    return (x, -y)


def twist(pt):
    if pt is None:
        return None
    _x, _y = pt
    # Field isomorphism from Z[p] / x**2 to Z[p] / x**2 - 18*x + 82
    xcoeffs = [_x.coeffs[0] - _x.coeffs[1] * 9, _x.coeffs[1]]
    ycoeffs = [_y.coeffs[0] - _y.coeffs[1] * 9, _y.coeffs[1]]
    # Isomorphism into subfield of Z[p] / w**12 - 18 * w**6 + 82, where w**6 = x
    nx = FQ12([int(xcoeffs[0])] + [0] * 5 + [int(xcoeffs[1])] + [0] * 5)
    ny = FQ12([int(ycoeffs[0])] + [0] * 5 + [int(ycoeffs[1])] + [0] * 5)
    # Divide x coord by w**2 and y coord by w**3
    return (nx * w ** 2, ny * w ** 3)


G12 = twist(G2)

from .pairing import pairing, final_exponentiate, miller_loop, cast_point_to_fq12, linefunc  # noqa: E402,F401
