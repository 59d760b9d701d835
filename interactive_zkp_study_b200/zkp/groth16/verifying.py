"""Drop-in for /root/reference/zkp/groth16/verifying.py (verifier-side group work).

The public-wire combination sum_i r_i * sigma1_3[i] (:34-37) is one GPU MSM; the four pairings and the
GT product stay on the CPU (O(1) per proof) and come from py_ecc -- where py_ecc is not installed a
caller passes `pairing=` (the tests hand in the oracle's restatement)."""
from ... import native
from ...compat import FQ, FR, G1, G2, HAVE_PY_ECC, curve_order, g1_from_ints  # noqa: F401

g1 = G1
g2 = G2


def _pairing_fn(pairing):
    if pairing is not None:
        return pairing
    if HAVE_PY_ECC:  # pragma: no cover
        from py_ecc import bn128
        return bn128.pairing
    raise NotImplementedError("pairings are outside the GPU hot path; install py_ecc or pass pairing=")


def _public_term(sigma1_3, rx_pub):
    pts = [sigma1_3[i] for i, _ in rx_pub]
    sc = [int(ri) % curve_order for _, ri in rx_pub]
    if not pts:
        return None
    return g1_from_ints(native.g1_msm(native.g1_vec_bytes(pts), native.fr_vec_bytes(sc), len(pts)))


def lhs(prf_A, prf_B, pairing=None):
    return _pairing_fn(pairing)(prf_B, prf_A)


def rhs(prf_C, sigma1_1, sigma1_3, sigma2_1, rx_pub, pairing=None):
    e = _pairing_fn(pairing)
    return (e(sigma2_1[0], sigma1_1[0]) * e(sigma2_1[1], _public_term(sigma1_3, rx_pub))) * e(sigma2_1[2], prf_C)


def verify(prf_A, prf_B, prf_C, sigma1_1, sigma1_3, sigma2_1, rx_pub, pairing=None):
    """e(A, B) == e(alpha, beta) * e(sum r_i sigma1_3[i], gamma) * e(C, delta)   (reference :29-40);
    rx_pub = [(index_i, r_i), ...]."""
    return lhs(prf_A, prf_B, pairing) == rhs(prf_C, sigma1_1, sigma1_3, sigma2_1, rx_pub, pairing)
