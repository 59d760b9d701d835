"""Drop-in for /root/reference/zkp/plonk/field.py: FR, curve constants, ec_* wrappers, roots of unity.

ec_mul / ec_add are the reference's only doors into py_ecc's group law (:72-103); here both are
tiny GPU MSMs (n = 1 and n = 2), so every group operation of the prover path -- also the ones outside
kzg.commit -- runs on the device and returns the same unique affine point.
"""
from ... import native
from ...compat import FQ, FQ2, FR, G1, G2, HAVE_PY_ECC, curve_order, g1_from_ints, g2_from_ints, is_g2  # noqa: F401

CURVE_ORDER = curve_order
Z1 = None


def ec_mul(point, scalar):
    """scalar * point (reference :72-88; the scalar is reduced mod the curve order)."""
    if isinstance(scalar, FR):
        scalar = int(scalar)
    k = int(scalar) % CURVE_ORDER
    if point is None:
        return None
    if is_g2(point):
        return g2_from_ints(native.g2_msm(native.g2_bytes(point), native.fe_bytes(k), 1))
    return g1_from_ints(native.g1_msm(native.g1_bytes(point), native.fe_bytes(k), 1))


def ec_add(p1, p2):
    """p1 + p2 (reference :91-103), including None operands, doubling and inverse points."""
    if p1 is None or p2 is None:
        return p1 if p2 is None else p2
    one = native.fe_bytes(1)
    if is_g2(p1):
        return g2_from_ints(native.g2_msm(native.g2_bytes(p1) + native.g2_bytes(p2), one + one, 2))
    return g1_from_ints(native.g1_msm(native.g1_bytes(p1) + native.g1_bytes(p2), one + one, 2))


def ec_neg(point):
    """-point (reference :106-115)."""
    if point is None:
        return None
    x, y = point
    return (x, -y)


def ec_pairing(g2_point, g1_point):
    """Pairings are verifier-side O(1) work and stay on the CPU in the reference (SURVEY.md 2,
    verifier rows).  Delegated to py_ecc where it is installed."""
    if HAVE_PY_ECC:  # pragma: no cover
        from py_ecc import bn128
        return bn128.pairing(g2_point, g1_point)
    raise NotImplementedError("ec_pairing is outside the GPU hot path; it needs py_ecc (verifier side)")


def get_root_of_unity(n):
    """Primitive n-th root of unity 5^((r-1)/n) (reference :145-182), same errors for bad n."""
    if n < 1 or (n & (n - 1)) != 0:
        raise ValueError(f"n은 2의 거듭제곱이어야 합니다: {n}")
    if n > (1 << 28):
        raise ValueError(f"n은 2^28 이하여야 합니다: {n}")
    if n == 1:
        return FR(1)
    return FR(5) ** ((CURVE_ORDER - 1) // n)


def get_roots_of_unity(n):
    """[1, w, w^2, ..., w^(n-1)] (reference :185-209) as one device NTT: the transform of the unit
    vector e_1 is exactly the list of powers of w."""
    omega = get_root_of_unity(n)
    if n == 1:
        return [FR(1)]
    e1 = [0, 1] + [0] * (n - 2)
    out = native.fr_ntt(native.fr_vec_bytes(e1), n.bit_length() - 1, int(omega))
    return [FR(v) for v in native.fr_vec_from_bytes(out)]
