"""Drop-in for the hot part of /root/reference/zkp/groth16/poly_utils.py.

Same names, argument order and return shapes; the Fr loops run on the GPU:
  _multiply_vec_matrix (:52-59)  -> zkp_fr_vec_matrix
  _multiply_polys (:17-22)       -> zkp_fr_poly_mul
  _add_polys/_subtract_polys     -> zkp_fr_vec_op
  _div_polys (:37-45)            -> zkp_fr_poly_divmod
  _eval_poly (:48-49)            -> zkp_fr_poly_eval
  hxr (:116-125)                 -> zkp_fr_vec_matrix x3 + zkp_groth16_quotient
including the reference's length quirk: R.Ax has len(R) = numWires entries of which the first
numGates are filled, so Hx has 2*numWires - numGates - 1 entries.
"""
from ... import native
from ...compat import FQ, FR, curve_order  # noqa: F401

_enc = native.fr_vec_bytes


def _fr_list(b):
    return [FR(v) for v in native.fr_vec_from_bytes(b)]


def _ints(v):
    return [int(x) % curve_order for x in v]


def _multiply_polys(a, b):
    a, b = _ints(a), _ints(b)
    return _fr_list(native.fr_poly_mul(_enc(a), len(a), _enc(b), len(b)))


def _add_polys(a, b, subtract=False):
    n = max(len(a), len(b))
    a = _ints(a) + [0] * (n - len(a))
    b = _ints(b) + [0] * (n - len(b))
    return _fr_list(native.fr_vec_op(1 if subtract else 0, _enc(a), _enc(b), n))


def _subtract_polys(a, b):
    return _add_polys(a, b, subtract=True)


def _div_polys(a, b):
    a, b = _ints(a), _ints(b)
    if len(a) < len(b):
        return [], [FR(v) for v in a]  # the reference's loop body never runs
    q, r = native.fr_poly_divmod(_enc(a), len(a), _enc(b), len(b))
    return _fr_list(q), _fr_list(r)


def _eval_poly(poly, x):
    p = _ints(poly)
    return FR(native.fr_poly_eval(_enc(p), len(p), int(x) % curve_order))


def _multiply_vec_matrix(vec, matrix):
    # len(vec) == number of rows; the result has len(vec) entries (reference :55), first len(row) filled
    assert not len(vec) == len(matrix[0])
    rows, cols = len(matrix), len(matrix[0])
    flat = [int(x) % curve_order for row in matrix for x in row]
    out = native.fr_vec_from_bytes(native.fr_vec_matrix(_enc(_ints(vec)), _enc(flat), rows, cols))
    target = [FR(0)] * len(vec)
    for j in range(cols):
        target[j] = FR(out[j])
    return target


def _multiply_vec_vec(vec1, vec2):
    assert len(vec1) == len(vec2)
    a, b = _ints(vec1), _ints(vec2)
    prod = native.fr_vec_op(2, _enc(a), _enc(b), len(a))
    # sum = evaluation of the product vector (as coefficients) at x = 1
    return FR(native.fr_poly_eval(prod, len(a), 1))


def getNumWires(Ax):
    return len(Ax)


def getNumGates(Ax):
    return len(Ax[0])


def getFRPoly1D(poly):
    return [FR(round(num)) for num in poly]


def getFRPoly2D(poly):
    return [[FR(round(num)) for num in vec] for vec in poly]


def _eval_rows(M, x_val):
    return [_eval_poly(row, x_val) for row in M]


ax_val = _eval_rows
bx_val = _eval_rows
cx_val = _eval_rows


def zx_val(Zx, x_val):
    return _eval_poly(Zx, x_val)


def hx_val(Hx, x_val):
    return _eval_poly(Hx, x_val)


def hxr(Ax, Bx, Cx, Zx, R):
    """(Ax.R * Bx.R - Cx.R) / Zx = Hx ... r      (reference :116-125)"""
    Rax = _multiply_vec_matrix(R, Ax)
    Rbx = _multiply_vec_matrix(R, Bx)
    Rcx = _multiply_vec_matrix(R, Cx)
    z = _ints(Zx)
    m = len(Rax)
    if 2 * m - 1 < len(z):  # the reference's division loop never runs: quotient of zeros, remainder = P
        Px = _subtract_polys(_multiply_polys(Rax, Rbx), Rcx)
        return _div_polys(Px, Zx)
    h, r = native.groth16_quotient(_enc(_ints(Rax)), _enc(_ints(Rbx)), _enc(_ints(Rcx)), m, _enc(z), len(z))
    return _fr_list(h), _fr_list(r)
