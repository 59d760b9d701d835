"""TEST INFRASTRUCTURE (oracle) -- optimal-Ate pairing on BN254 (``py_ecc.bn128.bn128_pairing``).

Used only as the acceptance check of the reference's verifiers
(/root/reference/zkp/groth16/verifying.py:29-40,
/root/reference/zkp/plonk/kzg.py:117-160, zkp/plonk/verifier.py).  Never on the
product path (pairings are O(1) per proof and stay on the CPU in the reference).
"""

from . import (
    FQ,
    FQ12,
    add,
    b,
    b2,
    curve_order,
    double,
    field_modulus,
    is_on_curve,
    twist,
)

ate_loop_count = 29793968203157093288
log_ate_loop_count = 63


def linefunc(P1, P2, T):
    """Line through P1, P2 evaluated at T."""
    assert P1 and P2 and T  # no infinity points here
    x1, y1 = P1
    x2, y2 = P2
    xt, yt = T
    if x1 != x2:
        m = (y2 - y1) / (x2 - x1)
        return m * (xt - x1) - (yt - y1)
    elif y1 == y2:
        m = 3 * x1 ** 2 / (2 * y1)
        return m * (xt - x1) - (yt - y1)
    else:
        return xt - x1


def cast_point_to_fq12(pt):
    if pt is None:
        return None
    x, y = pt
    return (FQ12([x.n] + [0] * 11), FQ12([y.n] + [0] * 11))


def miller_loop(Q, P):
    if Q is None or P is None:
        return FQ12.one()
    R = Q
    f = FQ12.one()
    for i in range(log_ate_loop_count, -1, -1):
        f = f * f * linefunc(R, R, P)
        R = double(R)
        if ate_loop_count & (2 ** i):
            f = f * linefunc(R, Q, P)
            R = add(R, Q)
    Q1 = (Q[0] ** field_modulus, Q[1] ** field_modulus)
    nQ2 = (Q1[0] ** field_modulus, -Q1[1] ** field_modulus)
    f = f * linefunc(R, Q1, P)
    R = add(R, Q1)
    f = f * linefunc(R, nQ2, P)
    return f ** ((field_modulus ** 12 - 1) // curve_order)


def pairing(Q, P):
    assert is_on_curve(Q, b2)
    assert is_on_curve(P, b)
    return miller_loop(twist(Q), cast_point_to_fq12(P))


def final_exponentiate(p):
    return p ** ((field_modulus ** 12 - 1) // curve_order)
