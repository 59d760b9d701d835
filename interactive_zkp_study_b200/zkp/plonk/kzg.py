"""Drop-in for /root/reference/zkp/plonk/kzg.py.

commit (:32-67) is the G1 MSM sum_i c_i [tau^i]_1: one GPU Pippenger call against a device-resident
copy of ``srs.g1_powers`` (cached by identity; the SRS is static).  Zero coefficients contribute
nothing, the all-zero polynomial commits to the point at infinity (``None``), and a polynomial
longer than the SRS raises the reference's ValueError.
"""
from ... import native, tables
from ...compat import g1_from_ints
from .field import FR, G1, CURVE_ORDER, ec_mul, ec_add, ec_neg, ec_pairing
from .polynomial import Polynomial, poly_div


def commit(poly, srs):
    if poly.degree > srs.max_degree:
        raise ValueError(
            f"다항식 차수 {poly.degree}가 SRS 최대 차수 {srs.max_degree}를 초과합니다"
        )
    coeffs = [int(c) % CURVE_ORDER for c in poly.coeffs]
    table = tables.g1_table(srs.g1_powers)
    return g1_from_ints(native.g1_msm_table(table, 0, native.fr_vec_bytes(coeffs), len(coeffs)))


def create_witness(poly, point, srs):
    """Opening proof pi = commit((p(x) - p(z)) / (x - z)) (reference :70-114)."""
    if not isinstance(point, FR):
        point = FR(point)
    y = poly.evaluate(point)
    quotient, remainder = poly_div(poly - Polynomial([y]), Polynomial([FR(0) - point, FR(1)]))
    for c in remainder.coeffs:
        if c != FR(0):
            raise ValueError("열기 증명 생성 실패: 나머지가 0이 아닙니다")
    return commit(quotient, srs)


def verify_opening(commitment, proof, point, evaluation, srs):
    """Pairing check e(C - y G1, G2) == e(pi, [tau - z]_2) (reference :117-160); verifier side,
    the group arithmetic goes through the GPU wrappers, the two pairings need py_ecc."""
    if not isinstance(point, FR):
        point = FR(point)
    if not isinstance(evaluation, FR):
        evaluation = FR(evaluation)
    tau_minus_z_g2 = ec_add(srs.g2_powers[1], ec_neg(ec_mul(srs.g2_powers[0], point)))
    c_minus_y = ec_add(commitment, ec_neg(ec_mul(G1, evaluation)))
    return ec_pairing(srs.g2_powers[0], c_minus_y) == ec_pairing(tau_minus_z_g2, proof)
