"""Drop-in for /root/reference/zkp/plonk/utils.py (coset FFTs :145-205 on the GPU; the scalar
helpers :25-116 keep their formulas)."""
from .field import FR
from .polynomial import Polynomial, fft, ifft, _ntt  # noqa: F401


def vanishing_poly_eval(n, zeta):
    return zeta ** n - FR(1)


def lagrange_basis_eval(i, n, omega, zeta):
    if not isinstance(zeta, FR):
        zeta = FR(zeta)
    omega_i = omega ** i
    zh_zeta = vanishing_poly_eval(n, zeta)
    denominator = zeta - omega_i
    if denominator == FR(0):
        return FR(1)
    n_inv = FR(1) / FR(n)
    return n_inv * zh_zeta * omega_i / denominator


def public_input_polynomial(pub_inputs, n, omega):
    if not pub_inputs:
        return Polynomial.zero()
    evals = [FR(0)] * n
    for i, val in enumerate(pub_inputs):
        evals[i] = val if isinstance(val, FR) else FR(val)
    return Polynomial.from_evaluations(evals, omega)


def public_input_poly_eval(pub_inputs, n, omega, zeta):
    result = FR(0)
    for i, val in enumerate(pub_inputs):
        if not isinstance(val, FR):
            val = FR(val)
        result = result + val * lagrange_basis_eval(i, n, omega, zeta)
    return result


def coset_fft(coeffs, omega, k=None):
    """Evaluate on the coset k*H: c_i <- c_i k^i fused into the transform's first pass."""
    if k is None:
        k = FR(5)
    return _ntt(coeffs, omega, False, shift=k)


def coset_ifft(evals, omega, k=None):
    """Inverse of coset_fft: inverse transform with c_i <- c_i k^-i fused into its last pass."""
    if k is None:
        k = FR(5)
    return _ntt(evals, omega, True, shift=k)


def pad_to_power_of_2(lst, fill=None):
    if fill is None:
        fill = FR(0)
    return list(lst) + [fill] * (next_power_of_2(len(lst)) - len(lst))


def next_power_of_2(n):
    if n <= 1:
        return 1
    p = 1
    while p < n:
        p <<= 1
    return p
