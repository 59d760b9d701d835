"""Drop-in for /root/reference/zkp/groth16/setup.py (CRS generation).

Every sigma list is "scalar_i * generator" for a vector of Fr scalars; the reference performs them as
sequential Python scalar multiplications (:15-69).  Here the scalars (powers of x, the
(beta*A_i + alpha*B_i + C_i)/gamma combinations, x^i*Z(x)/delta) are formed with device vector ops
and each list is ONE batched fixed-base multiplication kernel.  Masked wires keep the reference's
placeholder ``(FQ(0), FQ(0))`` (:39,:50), which is not a curve point and is never fed to the GPU.
"""
from ... import native
from ...compat import FQ, FR, G1, G2, curve_order, g1_from_ints, g2_from_ints

g1 = G1
g2 = G2

_enc = native.fr_vec_bytes
_dec = native.fr_vec_from_bytes


def _ints(v):
    return [int(x) % curve_order for x in v]


def _g1_batch(scalars):
    n = len(scalars)
    if n == 0:
        return []
    h = native.g1_fixed_base_mul(native.g1_bytes(G1), _enc(_ints(scalars)), n)
    raw = native.table_download(h, 0, n)
    h.free()
    return [g1_from_ints(native.g1_from_bytes(raw[64 * i:64 * i + 64])) for i in range(n)]


def _g2_batch(scalars):
    n = len(scalars)
    if n == 0:
        return []
    h = native.g2_fixed_base_mul(native.g2_bytes(G2), _enc(_ints(scalars)), n)
    raw = native.table_download(h, 0, n)
    h.free()
    return [g2_from_ints(native.g2_from_bytes(raw[128 * i:128 * i + 128])) for i in range(n)]


def _powers(x_val, count):
    """[1, x, ..., x^(count-1)] as ints (device prefix product)."""
    if count <= 0:
        return []
    return _dec(native.fr_prefix_product(_enc([int(x_val) % curve_order] * count), count))


def sigma11(alpha, beta, delta):
    return _g1_batch([alpha, beta, delta])


def sigma12(numGates, x_val):
    return _g1_batch(_powers(x_val, numGates))


def _abc_over(numWires, alpha, beta, Ax_val, Bx_val, Cx_val, denom):
    """(beta*A_i + alpha*B_i + C_i) / denom for every wire, on the device."""
    n = numWires
    a, b, c = _enc(_ints(Ax_val[:n])), _enc(_ints(Bx_val[:n])), _enc(_ints(Cx_val[:n]))
    t = native.fr_vec_op(0, native.fr_vec_op(3, a, native.fe_bytes(int(beta) % curve_order), n),
                         native.fr_vec_op(3, b, native.fe_bytes(int(alpha) % curve_order), n), n)
    t = native.fr_vec_op(0, t, c, n)
    inv = _dec(native.fr_batch_inverse(native.fe_bytes(int(denom) % curve_order), 1))[0]
    return _dec(native.fr_vec_op(3, t, native.fe_bytes(inv), n))


def sigma13(numWires, alpha, beta, gamma, Ax_val, Bx_val, Cx_val, pub_r_indexs=None):
    if pub_r_indexs == None:  # noqa: E711
        pub_r_indexs = [0, 1]
    print("sigma13 pub_r_indexs = {}".format(pub_r_indexs))
    vals = _abc_over(numWires, alpha, beta, Ax_val, Bx_val, Cx_val, gamma)
    VAL = [FR(0)] * numWires
    idx = [i for i in range(numWires) if i in pub_r_indexs]
    pts = _g1_batch([vals[i] for i in idx])
    sigma1_3 = [(FQ(0), FQ(0))] * numWires
    for i, p in zip(idx, pts):
        VAL[i] = FR(vals[i])
        sigma1_3[i] = p
    return sigma1_3, VAL


def sigma14(numWires, alpha, beta, delta, Ax_val, Bx_val, Cx_val, pub_r_indexs=None):
    if pub_r_indexs == None:  # noqa: E711
        pub_r_indexs = [0, 1]
    vals = _abc_over(numWires, alpha, beta, Ax_val, Bx_val, Cx_val, delta)
    idx = [i for i in range(numWires) if i not in pub_r_indexs]
    pts = _g1_batch([vals[i] for i in idx])
    sigma1_4 = [(FQ(0), FQ(0))] * numWires
    for i, p in zip(idx, pts):
        sigma1_4[i] = p
    return sigma1_4


def sigma15(numGates, delta, x_val, Zx_val):
    count = numGates - 1
    if count <= 0:
        return []
    inv_delta = _dec(native.fr_batch_inverse(native.fe_bytes(int(delta) % curve_order), 1))[0]
    k = int(Zx_val) % curve_order * inv_delta % curve_order
    scalars = _dec(native.fr_vec_op(3, _enc(_powers(x_val, count)), native.fe_bytes(k), count))
    return _g1_batch(scalars)


def sigma21(beta, delta, gamma):
    return _g2_batch([beta, gamma, delta])


def sigma22(numGates, x_val):
    return _g2_batch(_powers(x_val, numGates))
