"""Drop-in for /root/reference/zkp/plonk/verifier.py:42-208.

The reference rebuilds the transcript, then spends ~18 scalar multiplications forming the two G1
points of the final pairing check.  Here the same scalars are computed on the host and the two points
are two GPU MSMs (2 and 18 terms):

    A = W_zeta + u W_zeta_omega
    B = zeta W_zeta + u zeta omega W_zeta_omega + F + u [z] - E
    F = [t_lo] + zeta^n [t_mid] + zeta^2n [t_hi] + v D + v r0 G1 + v^2 [a] + v^3 [b] + v^4 [c] + v^5 [s1] + v^6 [s2]
    D = (a b)[q_m] + a [q_l] + b [q_r] + c [q_o] + [q_c] + (perm_z + alpha^2 L1) [z] - perm_s3 [s3]

and the check is e(A, [tau]_2) == e(B, [1]_2): the two pairings need py_ecc (or a `pairing=` callable).
"""
from ... import native
from ...compat import HAVE_PY_ECC, g1_from_ints
from .field import FR, G1, CURVE_ORDER
from .transcript import Transcript
from .permutation import K1, K2
from .utils import vanishing_poly_eval, lagrange_basis_eval


def _msm(terms):
    pts = [p for p, _ in terms]
    sc = [int(s) % CURVE_ORDER for _, s in terms]
    return g1_from_ints(native.g1_msm(native.g1_vec_bytes(pts), native.fr_vec_bytes(sc), len(pts)))


def verify(proof, public_inputs, preprocessed, srs, pairing=None):
    pp = preprocessed
    n, omega = pp.n, pp.omega
    t = Transcript()
    for name in ("a_comm", "b_comm", "c_comm"):
        t.append_point(name.encode(), getattr(proof, name))
    beta = t.challenge_scalar(b"beta")
    gamma = t.challenge_scalar(b"gamma")
    t.append_point(b"z_comm", proof.z_comm)
    alpha = t.challenge_scalar(b"alpha")
    for name in ("t_lo_comm", "t_mid_comm", "t_hi_comm"):
        t.append_point(name.encode(), getattr(proof, name))
    zeta = t.challenge_scalar(b"zeta")
    for name in ("a_eval", "b_eval", "c_eval", "s_sigma1_eval", "s_sigma2_eval", "z_omega_eval"):
        t.append_scalar(name.encode(), getattr(proof, name))
    v = t.challenge_scalar(b"v")
    u = t.challenge_scalar(b"u")

    a, b, c = proof.a_eval, proof.b_eval, proof.c_eval
    s1, s2, zw = proof.s_sigma1_eval, proof.s_sigma2_eval, proof.z_omega_eval
    zh = vanishing_poly_eval(n, zeta)
    l1 = lagrange_basis_eval(0, n, omega, zeta)
    perm_z = alpha * (a + beta * zeta + gamma) * (b + beta * K1 * zeta + gamma) * (c + beta * K2 * zeta + gamma)
    ab = (a + beta * s1 + gamma) * (b + beta * s2 + gamma)
    perm_s3 = alpha * ab * beta * zw
    r0 = FR(0) - alpha * ab * zw * (c + gamma) - alpha * alpha * l1     # PI(zeta) = 0 in the reference (:99)
    zeta_n = zeta ** n
    e_scalar = proof.r_eval / zh + v * proof.r_eval
    vp = v
    for val in (a, b, c, s1, s2):
        vp = vp * v
        e_scalar = e_scalar + vp * val
    e_scalar = e_scalar + u * zw

    A = _msm([(proof.W_zeta_comm, FR(1)), (proof.W_zeta_omega_comm, u)])
    v2 = v * v
    v4 = v2 * v2
    B = _msm([
        (proof.W_zeta_comm, zeta), (proof.W_zeta_omega_comm, u * zeta * omega),
        (proof.t_lo_comm, FR(1)), (proof.t_mid_comm, zeta_n), (proof.t_hi_comm, zeta_n * zeta_n),
        (pp.q_m_comm, v * a * b), (pp.q_l_comm, v * a), (pp.q_r_comm, v * b), (pp.q_o_comm, v * c), (pp.q_c_comm, v),
        (proof.z_comm, v * (perm_z + alpha * alpha * l1) + u), (pp.s_sigma3_comm, FR(0) - v * perm_s3),
        (G1, v * r0 - e_scalar),
        (proof.a_comm, v2), (proof.b_comm, v2 * v), (proof.c_comm, v4),
        (pp.s_sigma1_comm, v4 * v), (pp.s_sigma2_comm, v4 * v2),
    ])
    if pairing is None:
        if not HAVE_PY_ECC:
            raise NotImplementedError("pairings are outside the GPU hot path; install py_ecc or pass pairing=")
        from py_ecc import bn128  # pragma: no cover
        pairing = bn128.pairing   # pragma: no cover
    return pairing(srs.g2_powers[1], A) == pairing(srs.g2_powers[0], B)
