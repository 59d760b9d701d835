"""Round 5 (/root/reference/zkp/plonk/prover/round5.py:42-175): challenge v, the linearisation
polynomial r(x), the two opening quotients W_zeta and W_zeta_omega, two commitments.  The ~12
scalar*polynomial combinations, the evaluations and the two divisions by a linear factor run on the
GPU through Polynomial / poly_div."""
from ..field import FR
from ..polynomial import Polynomial, poly_div
from ..kzg import commit
from ..permutation import K1, K2
from ..utils import lagrange_basis_eval


def execute(state):
    state.v = state.transcript.challenge_scalar(b"v")
    v, n, zeta, omega = state.v, state.n, state.zeta, state.omega
    alpha, beta, gamma = state.alpha, state.beta, state.gamma
    pp, pr = state.preprocessed, state.proof
    a_e, b_e, c_e = pr.a_eval, pr.b_eval, pr.c_eval
    s1_e, s2_e, zw_e = pr.s_sigma1_eval, pr.s_sigma2_eval, pr.z_omega_eval

    pi_zeta = state.pi_poly.evaluate(zeta)
    l1_zeta = lagrange_basis_eval(0, n, omega, zeta)

    perm_z_scalar = alpha * (a_e + beta * zeta + gamma) * (b_e + beta * K1 * zeta + gamma) * (c_e + beta * K2 * zeta + gamma)
    ab_factor = (a_e + beta * s1_e + gamma) * (b_e + beta * s2_e + gamma)
    perm_s3_scalar = alpha * ab_factor * beta * zw_e
    perm_const = FR(0) - alpha * ab_factor * zw_e * (c_e + gamma)
    aa_l1 = alpha * alpha * l1_zeta

    r_poly = (pp.q_m_poly * (a_e * b_e) + pp.q_l_poly * a_e + pp.q_r_poly * b_e + pp.q_o_poly * c_e
              + pp.q_c_poly + Polynomial([pi_zeta]))
    r_poly = r_poly + state.z_poly * perm_z_scalar
    r_poly = r_poly - pp.s_sigma3_poly * perm_s3_scalar
    r_poly = r_poly + Polynomial([perm_const])
    r_poly = r_poly + state.z_poly * aa_l1
    r_poly = r_poly + Polynomial([FR(0) - aa_l1])
    r_eval = r_poly.evaluate(zeta)
    pr.r_eval = r_eval

    zeta_n = zeta ** n
    zeta_2n = zeta_n * zeta_n
    t_eval = (state.t_lo_poly.evaluate(zeta) + zeta_n * state.t_mid_poly.evaluate(zeta)
              + zeta_2n * state.t_hi_poly.evaluate(zeta))
    num = state.t_lo_poly + state.t_mid_poly * zeta_n + state.t_hi_poly * zeta_2n - Polynomial([t_eval])
    num = num + (r_poly - Polynomial([r_eval])) * v
    v_pow = v
    for poly, ev in ((state.a_poly, a_e), (state.b_poly, b_e), (state.c_poly, c_e),
                     (pp.s_sigma1_poly, s1_e), (pp.s_sigma2_poly, s2_e)):
        v_pow = v_pow * v
        num = num + (poly - Polynomial([ev])) * v_pow

    W_zeta, _ = poly_div(num, Polynomial([FR(0) - zeta, FR(1)]))
    W_zeta_omega, _ = poly_div(state.z_poly - Polynomial([zw_e]), Polynomial([FR(0) - zeta * omega, FR(1)]))
    pr.W_zeta_comm = commit(W_zeta, state.srs)
    pr.W_zeta_omega_comm = commit(W_zeta_omega, state.srs)
