"""Drop-in for /root/reference/zkp/plonk/utils.py: same callables, same values.

The vector work (coset transforms :145-205, the public-input interpolation :84-116) runs on the GPU;
the closed-form scalar helpers (:25-81, :119-142, :208-246) are single field expressions evaluated on
the host, as in the reference.
"""
from .field import FR
from .polynomial import Polynomial, fft, ifft, _ntt  # noqa: F401

DEFAULT_COSET_SHIFT = 5  # the reference's conventional generator (utils.py:166-167)


def _fr(x):
    return x if isinstance(x, FR) else FR(x)


def vanishing_poly_eval(n, zeta):
    """Z_H(zeta) = zeta^n - 1."""
    return zeta ** n - FR(1)


def lagrange_basis_eval(i, n, omega, zeta):
    """L_i(zeta) = w^i (zeta^n - 1) / (n (zeta - w^i)); 1 when zeta is the i-th domain point."""
    zeta = _fr(zeta)
    w_i = omega ** i
    gap = zeta - w_i
    if gap == FR(0):
        return FR(1)
    return (FR(1) / FR(n)) * vanishing_poly_eval(n, zeta) * w_i / gap


def public_input_polynomial(pub_inputs, n, omega):
    """PI(x): the polynomial whose first len(pub_inputs) evaluations on H are the public inputs and the
    rest zero -- one device iNTT."""
    if not pub_inputs:
        return Polynomial.zero()
    values = [_fr(v) for v in pub_inputs]
    return Polynomial.from_evaluations(values + [FR(0)] * (n - len(values)), omega)


def public_input_poly_eval(pub_inputs, n, omega, zeta):
    """PI(zeta) = sum_i w_i L_i(zeta) without building the polynomial."""
    total = FR(0)
    for i, value in enumerate(pub_inputs):
        total = total + _fr(value) * lagrange_basis_eval(i, n, omega, zeta)
    return total


def coset_fft(coeffs, omega, k=None):
    """Evaluations on the coset k*H: the scaling c_i <- c_i k^i is fused into the transform's first pass."""
    return _ntt(coeffs, omega, False, shift=FR(DEFAULT_COSET_SHIFT) if k is None else k)


def coset_ifft(evals, omega, k=None):
    """Inverse of coset_fft: inverse transform with c_i <- c_i k^-i fused into its last pass."""
    return _ntt(evals, omega, True, shift=FR(DEFAULT_COSET_SHIFT) if k is None else k)


def next_power_of_2(n):
    return 1 if n <= 1 else 1 << (n - 1).bit_length()


def pad_to_power_of_2(lst, fill=None):
    filler = FR(0) if fill is None else fill
    return list(lst) + [filler] * (next_power_of_2(len(lst)) - len(lst))
