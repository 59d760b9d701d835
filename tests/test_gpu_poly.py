"""-m gpu: Fr vector / polynomial entry points against the oracle's restatement of the reference's
coefficient-form loops (polynomial.py:85-159,385-435; poly_utils.py:17-59,116-125)."""
import random

import pytest

from oracle import ref_path

pytestmark = pytest.mark.gpu
R = ref_path.R


def _rand(rng, n):
    return [rng.randrange(R) for _ in range(n)]


def test_vec_ops(native):
    rng = random.Random(1)
    n = 1000
    a, b = _rand(rng, n), _rand(rng, n)
    a[:4] = [0, 1, R - 1, R - 1]
    b[:4] = [0, R - 1, R - 1, 1]
    ab, bb = native.fr_vec_bytes(a), native.fr_vec_bytes(b)
    dec = native.fr_vec_from_bytes
    assert dec(native.fr_vec_op(0, ab, bb, n)) == [(x + y) % R for x, y in zip(a, b)]
    assert dec(native.fr_vec_op(1, ab, bb, n)) == [(x - y) % R for x, y in zip(a, b)]
    assert dec(native.fr_vec_op(2, ab, bb, n)) == [(x * y) % R for x, y in zip(a, b)]
    assert dec(native.fr_vec_op(3, ab, native.fe_bytes(b[7]), n)) == [(x * b[7]) % R for x in a]


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 1000, 5000])
def test_batch_inverse(native, n):
    rng = random.Random(n)
    a = _rand(rng, n)
    if n > 2:
        a[1] = 0
        a[n - 1] = 0
    got = native.fr_vec_from_bytes(native.fr_batch_inverse(native.fr_vec_bytes(a), n))
    assert got == [ref_path.inv(x) for x in a]


@pytest.mark.parametrize("n", [1, 2, 63, 64, 65, 1000, 70000])
def test_poly_eval_is_horner(native, n):
    rng = random.Random(n)
    c = _rand(rng, n)
    for x in (0, 1, rng.randrange(R)):
        assert native.fr_poly_eval(native.fr_vec_bytes(c), n, x) == ref_path.poly_eval(c, x)


@pytest.mark.parametrize("la,lb", [(1, 1), (1, 5), (2, 2), (3, 6), (17, 33), (100, 129), (700, 300)])
def test_poly_mul_matches_schoolbook(native, la, lb):
    rng = random.Random(la * 1000 + lb)
    a, b = _rand(rng, la), _rand(rng, lb)
    want = [0] * (la + lb - 1)
    for i, x in enumerate(a):
        for j, y in enumerate(b):
            want[i + j] = (want[i + j] + x * y) % R
    got = native.fr_vec_from_bytes(native.fr_poly_mul(native.fr_vec_bytes(a), la, native.fr_vec_bytes(b), lb))
    assert got == want


def _pad(v, n):
    return list(v) + [0] * (n - len(v))


@pytest.mark.parametrize("la,lb", [(1, 1), (5, 1), (5, 2), (6, 6), (9, 4), (40, 13), (300, 129), (513, 2), (600, 300)])
def test_poly_divmod_matches_long_division(native, la, lb):
    rng = random.Random(la * 1000 + lb)
    a, b = _rand(rng, la), _rand(rng, lb)
    q, r = ref_path.poly_div(a, b)
    gq, gr = native.fr_poly_divmod(native.fr_vec_bytes(a), la, native.fr_vec_bytes(b), lb)
    assert native.fr_vec_from_bytes(gq) == _pad(q, la - lb + 1)
    assert native.fr_vec_from_bytes(gr) == (_pad(r, lb - 1) if lb > 1 else [])


def test_poly_divmod_exact_and_vanishing(native):
    rng = random.Random(77)
    n = 64
    zh = [R - 1] + [0] * (n - 1) + [1]           # x^n - 1 (Polynomial.vanishing, polynomial.py:227-236)
    t = _rand(rng, 3 * n + 6)
    c = ref_path.poly_mul(t, zh)
    gq, gr = native.fr_poly_divmod(native.fr_vec_bytes(c), len(c), native.fr_vec_bytes(zh), len(zh))
    assert native.fr_vec_from_bytes(gq) == t
    assert native.fr_vec_from_bytes(gr) == [0] * n
    c[5] = (c[5] + 1) % R                        # not divisible any more: remainder shows it
    q, r = ref_path.poly_div(c, zh)
    gq, gr = native.fr_poly_divmod(native.fr_vec_bytes(c), len(c), native.fr_vec_bytes(zh), len(zh))
    assert native.fr_vec_from_bytes(gq) == _pad(q, len(c) - len(zh) + 1)
    assert native.fr_vec_from_bytes(gr) == _pad(r, n)
    # division by a linear factor (round 5 / create_witness, kzg.py:95-104)
    p = _rand(rng, 200)
    z = rng.randrange(R)
    y = ref_path.poly_eval(p, z)
    pm = list(p)
    pm[0] = (pm[0] - y) % R
    q, r = ref_path.poly_div(pm, [(-z) % R, 1])
    gq, gr = native.fr_poly_divmod(native.fr_vec_bytes(pm), 200, native.fr_vec_bytes([(-z) % R, 1]), 2)
    assert native.fr_vec_from_bytes(gq) == _pad(q, 199) and native.fr_vec_from_bytes(gr) == [0]


def test_vec_matrix(native):
    rng = random.Random(3)
    rows, cols = 6, 4
    vec = _rand(rng, rows)
    mat = [_rand(rng, cols) for _ in range(rows)]
    flat = [x for row in mat for x in row]
    got = native.fr_vec_from_bytes(native.fr_vec_matrix(native.fr_vec_bytes(vec), native.fr_vec_bytes(flat), rows, cols))
    assert got == [sum(vec[i] * mat[i][j] for i in range(rows)) % R for j in range(cols)]


@pytest.mark.parametrize("m,k", [(6, 4), (20, 9), (150, 100)])
def test_groth16_quotient_matches_hxr(native, m, k):
    """hxr with the reference's length quirk: vectors of numWires entries, first numGates filled."""
    rng = random.Random(m)
    Ax = [_rand(rng, k) for _ in range(m)]
    Bx = [_rand(rng, k) for _ in range(m)]
    Cx = [_rand(rng, k) for _ in range(m)]
    Rv = _rand(rng, m)
    Z = [1]
    for i in range(1, k + 1):                     # Z(x) = prod (x - i), qap_creator_lcm.py:128-135
        Z = ref_path.g16_multiply_polys(Z, [(-i) % R, 1])
    Hx, rem = ref_path.hxr(Ax, Bx, Cx, Z, Rv)
    ra = ref_path.g16_multiply_vec_matrix(Rv, Ax)
    rb = ref_path.g16_multiply_vec_matrix(Rv, Bx)
    rc = ref_path.g16_multiply_vec_matrix(Rv, Cx)
    enc = native.fr_vec_bytes
    gh, gr = native.groth16_quotient(enc(ra), enc(rb), enc(rc), m, enc(Z), len(Z))
    assert native.fr_vec_from_bytes(gh) == Hx
    assert native.fr_vec_from_bytes(gr) == rem


def test_wire_format_moves_device_vectors_without_python_integers(native):
    """wire.serialize_handle / deserialize_to_handle: the bytes of a device-resident
    polynomial are exactly what the list-level serializer writes for the same coefficients."""
    import random
    from interactive_zkp_study_b200.zkp.plonk import wire
    from interactive_zkp_study_b200.zkp.plonk.polynomial import Polynomial
    from interactive_zkp_study_b200.zkp.plonk.field import FR
    rng = random.Random(31)
    coeffs = [rng.randrange(1, native.R_MOD) for _ in range(777)]
    h = native.scalars_load(native.fr_vec_bytes(coeffs), len(coeffs))
    blob = wire.serialize_handle(h)
    assert blob == wire.serialize_poly(Polynomial([FR(c) for c in coeffs]))
    assert wire.serialize_handle(h, 10, 5) == blob[32 * 5:32 * 15]
    h2 = wire.deserialize_to_handle(blob)
    assert h2.n == len(coeffs) and native.scalars_download(h2, 0, h2.n) == blob
    assert [int(c) for c in wire.deserialize_poly(blob).coeffs] == coeffs


def test_batched_evaluation_and_linear_combination(native):
    """zkp_fr_poly_eval_multi_dev / zkp_fr_lincomb_dev against Horner and Python sums (ragged lengths,
    two evaluation points, a zero coefficient, a single-coefficient polynomial)."""
    import random
    from oracle import ref_path
    rng = random.Random(77)
    Rm = native.R_MOD
    lens = [1, 5, 64, 65, 4097, 5000, 5003]
    polys = [[rng.randrange(Rm) for _ in range(n)] for n in lens]
    hs = [native.scalars_load(native.fr_vec_bytes(p), len(p)) for p in polys]
    x0, x1 = rng.randrange(Rm), rng.randrange(Rm)
    items = [(h, 0, len(p), x0 if i % 3 else x1) for i, (h, p) in enumerate(zip(hs, polys))]
    got = native.fr_poly_eval_multi_dev(items)
    assert got == [ref_path.poly_eval(p, x0 if i % 3 else x1) for i, p in enumerate(polys)]
    assert native.fr_poly_eval_multi_dev([(hs[5], 7, 100, x0)]) == [ref_path.poly_eval(polys[5][7:107], x0)]
    coeffs = [rng.randrange(Rm) for _ in lens]
    coeffs[2] = 0
    n = 5003
    dst = native.scalars_generate(9, n)                  # pre-filled: the call overwrites, it does not accumulate
    native.lincomb_dev(dst, 0, n, [(c, h, 0, len(p)) for c, h, p in zip(coeffs, hs, polys)])
    want = [sum(c * p[i] for c, p in zip(coeffs, polys) if i < len(p)) % Rm for i in range(n)]
    assert native.fr_vec_from_bytes(native.scalars_download(dst, 0, n)) == want
    with pytest.raises(native.ZkpB200Error):
        native.lincomb_dev(dst, 0, 10, [(1, hs[3], 0, 65)])          # item longer than the destination
    with pytest.raises(native.ZkpB200Error):
        native.lincomb_dev(hs[6], 0, 5003, [(1, hs[6], 0, 5003)])    # in-place is refused
