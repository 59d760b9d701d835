"""Round 3 (/root/reference/zkp/plonk/prover/round3.py:56-184): the quotient
t(x) = [gate + alpha*perm + alpha^2*(z-1)*L1] / (x^n - 1), split in three, three commitments.

The reference forms every product by O(n^2) schoolbook multiplication, builds L1 with an O(n^2)
Lagrange product and divides by long division.  Here every product is an NTT product on the GPU,
L1 is the iNTT of the unit vector (the same unique polynomial), z(w x) is a device power scaling,
and the division by x^n - 1 is the stride recurrence kernel.  A non-zero remainder raises the
reference's ValueError (:150-155)."""
from .... import native
from ..field import FR, CURVE_ORDER
from ..polynomial import Polynomial, poly_div
from ..kzg import commit
from ..permutation import K1, K2


def execute(state):
    state.alpha = state.transcript.challenge_scalar(b"alpha")
    n, omega = state.n, state.omega
    alpha, beta, gamma = state.alpha, state.beta, state.gamma
    pp = state.preprocessed
    a, b, c, z, pi = state.a_poly, state.b_poly, state.c_poly, state.z_poly, state.pi_poly

    # z(w x): c_i <- c_i w^i, powers of w from the device prefix-product scan
    zl = len(z.coeffs)
    w_pows = native.fr_prefix_product(native.fr_vec_bytes([int(omega)] * zl), zl)
    zc = native.fr_vec_bytes([int(v) % CURVE_ORDER for v in z.coeffs])
    z_omega = Polynomial._from_ints(native.fr_vec_from_bytes(native.fr_vec_op(2, zc, w_pows, zl)))

    x_poly = Polynomial([FR(0), FR(1)])
    l1 = Polynomial.from_evaluations([FR(1)] + [FR(0)] * (n - 1), omega)  # == lagrange_basis(domain, 0)
    g = Polynomial([gamma])

    term1 = pp.q_l_poly * a + pp.q_r_poly * b + pp.q_o_poly * c + pp.q_m_poly * (a * b) + pp.q_c_poly + pi
    perm_num = (a + x_poly * beta + g) * (b + x_poly * (beta * K1) + g) * (c + x_poly * (beta * K2) + g) * z
    perm_den = ((a + pp.s_sigma1_poly * beta + g) * (b + pp.s_sigma2_poly * beta + g)
                * (c + pp.s_sigma3_poly * beta + g) * z_omega)
    term2 = (perm_num - perm_den) * alpha
    term3 = (z - Polynomial([FR(1)])) * l1 * (alpha * alpha)
    constraint = term1 + term2 + term3

    t_poly, remainder = poly_div(constraint, Polynomial.vanishing(n))
    for coeff in remainder.coeffs:
        if coeff != FR(0):
            raise ValueError(
                "제약 다항식이 Z_H(x)로 나누어 떨어지지 않습니다. "
                "회로 또는 witness에 오류가 있습니다."
            )

    t = list(t_poly.coeffs)
    t += [FR(0)] * (3 * n - len(t))
    state.t_lo_poly = Polynomial(t[:n])
    state.t_mid_poly = Polynomial(t[n:2 * n])
    state.t_hi_poly = Polynomial(t[2 * n:])  # coefficients past 3n stay in t_hi (reference :169-171)
    for part in ("t_lo", "t_mid", "t_hi"):
        setattr(state.proof, part + "_comm", commit(getattr(state, part + "_poly"), state.srs))
    for part in ("t_lo", "t_mid", "t_hi"):
        state.transcript.append_point(part.encode() + b"_comm", getattr(state.proof, part + "_comm"))

