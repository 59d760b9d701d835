"""Drop-in for /root/reference/zkp/plonk/permutation.py.

compute_accumulator (:89-137) is n-1 sequential (multiply, divide) steps in the reference.  Here:
numerators and denominators are built with device vector ops, the n-1 divisions become ONE batch
inversion (Montgomery's trick on the GPU), and the running product z_{i+1} = z_i * ratio_i is
evaluated as prefix products.  Each z_i is a field element, so the values are the reference's.
"""
from ... import native
from .field import FR, CURVE_ORDER

K1 = FR(2)
K2 = FR(3)

_enc = native.fr_vec_bytes
_dec = native.fr_vec_from_bytes


def build_permutation_polynomials(sigma, n, domain):
    """sigma (3n positions) -> evaluations of S_sigma1..3 (reference :44-86): a gather from the table
    [H, K1*H, K2*H], built with two device scalings."""
    d = [int(x) % CURVE_ORDER for x in domain]
    db = _enc(d)
    table = d + _dec(native.fr_vec_op(3, db, native.fe_bytes(int(K1)), n)) + \
        _dec(native.fr_vec_op(3, db, native.fe_bytes(int(K2)), n))
    pick = [FR(table[sigma[i]]) for i in range(3 * n)]
    return pick[:n], pick[n:2 * n], pick[2 * n:]


def _affine(vals, scale, base, gamma_b, n):
    """vals[i] + scale * base[i] + gamma, all on the device."""
    t = native.fr_vec_op(3, base, native.fe_bytes(scale), n)
    t = native.fr_vec_op(0, t, vals, n)
    return native.fr_vec_op(0, t, gamma_b, n)


def compute_accumulator(a_vals, b_vals, c_vals, sigma, n, domain, beta, gamma):
    if n == 1:
        return [FR(1)]
    s1, s2, s3 = build_permutation_polynomials(sigma, n, domain)
    beta_i, gamma_i = int(beta) % CURVE_ORDER, int(gamma) % CURVE_ORDER
    ints = lambda v: _enc([int(x) % CURVE_ORDER for x in v])
    a, b, c, dom = ints(a_vals), ints(b_vals), ints(c_vals), ints(domain)
    gam = _enc([gamma_i] * n)
    k1b, k2b = beta_i * int(K1) % CURVE_ORDER, beta_i * int(K2) % CURVE_ORDER
    mul = lambda x, y: native.fr_vec_op(2, x, y, n)
    num = mul(mul(_affine(a, beta_i, dom, gam, n), _affine(b, k1b, dom, gam, n)), _affine(c, k2b, dom, gam, n))
    den = mul(mul(_affine(a, beta_i, ints(s1), gam, n), _affine(b, beta_i, ints(s2), gam, n)),
              _affine(c, beta_i, ints(s3), gam, n))
    ratio = mul(num, native.fr_batch_inverse(den, n))
    z = native.fr_prefix_product(ratio, n)  # z[0] = 1, z[i+1] = z[i] * ratio[i]
    return [FR(v) for v in _dec(z)]
