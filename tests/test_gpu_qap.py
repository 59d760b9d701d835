"""-m gpu: QAP construction at scale: sparse R1CS -> witness polynomials by sparse
mat-vec + exact interpolation on {1..k} (subproduct tree on batched NTTs) -> CRS -> proof.

Parity: (1) every stage against the reference's own R1CS -> QAP -> setup -> proof pipeline on the two
programs of tests/golden/groth16_qap.json (bit-identical proof); (2) against the oracle's exact
restatement of qap_creator_lcm at sizes the reference's floats cannot reach; (3) at 2^14 constraints
through closed-form discrete logs and the pairing equation in the exponent."""
import random

import pytest

from oracle import bn254, ref_path
from tests.util import g1, g2, ints, load

pytestmark = pytest.mark.gpu
R = bn254.R


def _dl(native, h, n, off=0):
    return native.fr_vec_from_bytes(native.scalars_download(h, off, n))


def _g1pts(native, table, off, n):
    raw = native.table_download(table, off, n)
    return [native.g1_from_bytes(raw[64 * i:64 * i + 64]) for i in range(n)]


def _pt(p):
    return None if p is None else (int(p[0]), int(p[1]))


def _pt2(p):
    return None if p is None else ((int(p[0].coeffs[0]), int(p[0].coeffs[1])), (int(p[1].coeffs[0]), int(p[1].coeffs[1])))


@pytest.mark.parametrize("case", [0, 1])
def test_golden_pipeline_bit_exact(native, case):
    from interactive_zkp_study_b200.zkp.groth16 import qap_device as qd
    c = load("groth16_qap.json")["cases"][case]
    k, m = c["numGates"], c["numWires"]
    r1cs = qd.SparseR1CS.from_dense(c["r1cs_A"], c["r1cs_B"], c["r1cs_C"])
    assert (r1cs.k, r1cs.m) == (k, m)
    dev = qd.DeviceR1CS(r1cs)
    w = native.scalars_load(native.fr_vec_bytes(c["witness"]), m)
    uA, uB, uC = qd.witness_polys(dev, w)
    assert _dl(native, uA, k) == ints(c["uA"])
    assert _dl(native, uB, k) == ints(c["uB"])
    assert _dl(native, uC, k) == ints(c["uC"])
    t = {key: int(v) for key, v in c["toxic"].items()}
    keys = qd.setup(dev, t["alpha"], t["beta"], t["gamma"], t["delta"], t["x_val"], pub_r_indexs=c["pub_r_indexs"],
                    precompute=False)
    assert _dl(native, keys.Z, k + 1) == ints(c["Zx"])
    key = keys.device_key
    # the CRS tables hold exactly the reference's sigma elements
    assert _g1pts(native, key.TA, 0, k) == [g1(p) for p in c["sigma1_2"]]
    assert _g1pts(native, key.TA, k, 2) == [g1(c["sigma1_1"][0]), g1(c["sigma1_1"][2])]
    priv = [i for i in range(m) if i not in c["pub_r_indexs"]]
    assert _g1pts(native, key.TC, k + 1, len(priv)) == [g1(c["sigma1_4"][i]) for i in priv]
    assert _g1pts(native, key.TC, k + 1 + len(priv), k - 1) == [g1(p) for p in c["sigma1_5"]]
    raw2 = native.table_download(key.TB2, 0, k)
    assert [native.g2_from_bytes(raw2[128 * i:128 * i + 128]) for i in range(k)] == [g2(p) for p in c["sigma2_2"]]
    assert [_pt(p) for p in keys.sigma1_1] == [g1(p) for p in c["sigma1_1"]]
    assert [_pt2(p) for p in keys.sigma2_1] == [g2(p) for p in c["sigma2_1"]]
    for i in c["pub_r_indexs"]:
        assert _pt(keys.sigma1_3[i]) == g1(c["sigma1_3"][i])
    for i in priv:
        assert _pt(keys.sigma1_3[i]) == (0, 0)          # the reference's placeholder (setup.py:37)
    A, B, C = qd.prove(keys, dev, w, int(c["r"]), int(c["s"]))
    assert _pt(A) == g1(c["proof_a"]) and _pt2(B) == g2(c["proof_b"]) and _pt(C) == g1(c["proof_c"])


@pytest.mark.parametrize("k", [1, 2, 3, 5, 8, 13, 31, 64, 100])
def test_interpolation_vanishing_lagrange_match_oracle(native, k):
    rng = random.Random(900 + k)
    y = [rng.randrange(R) for _ in range(k)]
    if k > 3:
        y[2] = 0
    h = native.scalars_load(native.fr_vec_bytes(y), k)
    out = native.scalars_alloc(k)
    native.fr_ap_interpolate_dev(h, k, out)
    want = ref_path.qap_lagrange_interp(y)
    assert _dl(native, out, k) == want
    scale = rng.randrange(1, R)
    native.fr_ap_interpolate_dev(h, k, out, scale=scale)
    assert _dl(native, out, k) == [c * scale % R for c in want]
    Z = [1]
    for i in range(1, k + 1):
        Z = ref_path.g16_multiply_polys(Z, [(-i) % R, 1])
    assert _dl(native, native.fr_ap_vanishing_dev(k), k + 1) == Z
    x = rng.randrange(R)
    lag = _dl(native, native.fr_ap_lagrange_dev(k, x), k)
    assert lag == [ref_path.poly_eval(ref_path.qap_mk_singleton(j + 1, 1, k), x) for j in range(k)]
    on = (k + 1) // 2                                   # x on the domain: a unit vector
    assert _dl(native, native.fr_ap_lagrange_dev(k, on), k) == [1 if j + 1 == on else 0 for j in range(k)]


@pytest.mark.parametrize("k", [1000, 1024, 1025, 5000])
def test_interpolation_round_trip_medium(native, k):
    """p(j + 1) == y_j at every point (Horner on the host) and deg p < k, non-power-of-two sizes included."""
    rng = random.Random(k)
    y = [rng.randrange(R) for _ in range(k)]
    h = native.scalars_load(native.fr_vec_bytes(y), k)
    out = native.scalars_alloc(k)
    native.fr_ap_interpolate_dev(h, k, out)
    coeffs = _dl(native, out, k)
    for j in rng.sample(range(k), 40) + [0, k - 1]:
        assert ref_path.poly_eval(coeffs, j + 1) == y[j]
    z = _dl(native, native.fr_ap_vanishing_dev(k), k + 1)
    assert z[k] == 1 and all(ref_path.poly_eval(z, j) == 0 for j in (1, 2, k // 2, k)) and ref_path.poly_eval(z, k + 1) != 0
    x = rng.randrange(R)
    lag = _dl(native, native.fr_ap_lagrange_dev(k, x), k)
    assert sum(a * b for a, b in zip(lag, y)) % R == ref_path.poly_eval(coeffs, x)


def test_sparse_matvec_and_errors(native):
    from interactive_zkp_study_b200.native import ZkpB200Error
    rng = random.Random(5)
    rows, cols = 200, 150
    dense = [[(rng.randrange(R) if rng.random() < 0.05 else 0) for _ in range(cols)] for _ in range(rows)]
    dense[7] = [0] * cols                               # empty row
    dense[9] = [rng.randrange(R) for _ in range(cols)]  # dense row
    rp, ci, vals = [0], [], []
    for row in dense:
        for j, v in enumerate(row):
            if v:
                ci.append(j)
                vals.append(v)
        rp.append(len(ci))
    mat = native.sparse_load(rp, ci, vals, rows, cols)
    vec = [rng.randrange(R) for _ in range(cols)]
    v = native.scalars_load(native.fr_vec_bytes(vec), cols)
    out = native.scalars_alloc(rows)
    native.sparse_matvec_dev(mat, v, out)
    assert _dl(native, out, rows) == [sum(a * b for a, b in zip(row, vec)) % R for row in dense]
    with pytest.raises(ZkpB200Error):
        native.sparse_load([0, 2, 1], [0, 1], [1, 2], 2, 4)           # row_ptr not monotone / wrong end
    with pytest.raises(ZkpB200Error):
        native.sparse_load([0, 1, 2], [0, 9], [1, 2], 2, 4)           # column out of range
    short = native.scalars_alloc(cols - 1)
    with pytest.raises(ZkpB200Error):
        native.sparse_matvec_dev(mat, short, out)                      # vector shorter than the column count
    with pytest.raises(ZkpB200Error):
        native.sparse_matvec_dev(v, v, out)                            # not a matrix handle
    empty = native.sparse_load([0], [], [], 0, 3)
    assert empty.n == 0


def _synthetic_circuit(k, seed):
    """k gates over m = k + 2 wires: wire 0 = 1, wire 1 = input, gate g defines wire g + 2 as a product of
    two earlier wires or as a linear combination of them (times 1)."""
    rng = random.Random(seed)
    w = [1, rng.randrange(2, 1 << 20)]
    ra, rb, rc = [], [], []
    for g in range(k):
        a, b = rng.randrange(len(w)), rng.randrange(len(w))
        if rng.random() < 0.5:
            ra.append({a: 1})
            rb.append({b: 1})
            w.append(w[a] * w[b] % R)
        else:
            cst = rng.randrange(1, 1 << 16)
            row = {a: 1}
            row[b] = (row.get(b, 0) + cst) % R
            ra.append(row)
            rb.append({0: 1})
            w.append((w[a] + cst * w[b]) % R)
        rc.append({g + 2: 1})
    return ra, rb, rc, w


@pytest.mark.parametrize("k,lcm,precompute", [(37, True, False), (300, False, False), (1 << 14, True, True)])
def test_synthetic_circuit_proof_satisfies_pairing_equation(native, k, lcm, precompute):
    """Sparse R1CS with a real witness: zero remainder, every proof element equals its closed-form
    discrete log times the generator, and the Groth16 verification equation holds in the exponent:
    a*b == alpha*beta + (sum_pub w_i val_i) + c*delta   (verifying.py:29-40 with e(g1, g2) factored out)."""
    from interactive_zkp_study_b200.zkp.groth16 import qap_device as qd
    m = k + 2
    ra, rb, rc, wit = _synthetic_circuit(k, 4242 + k)
    r1cs = qd.SparseR1CS.from_rows(k, m, ra, rb, rc)
    dev = qd.DeviceR1CS(r1cs)
    rng = random.Random(k)
    alpha, beta, gamma, delta, x = (rng.randrange(1, R) for _ in range(5))
    keys = qd.setup(dev, alpha, beta, gamma, delta, x, lcm=lcm, precompute=precompute)
    w = native.scalars_load(native.fr_vec_bytes(wit), m)
    r, s = rng.randrange(R), rng.randrange(R)
    A, B, C, hq, hr, uA, uB, uC = qd.prove(keys, dev, w, r, s, keep=True)
    assert native.scalars_is_zero(hr, 0, k)                              # the circuit is satisfied: Z | uA*uB - uC
    ev = lambda h, n: native.fr_poly_eval_dev(h, 0, n, x)
    ax, bx, cx, zx = ev(uA, k), ev(uB, k), ev(uC, k), ev(keys.Z, k + 1)
    hx = ev(hq, k - 1) if k > 1 else 0
    assert (ax * bx - cx) % R == hx * zx % R
    a = (alpha + ax + r * delta) % R
    b = (beta + bx + s * delta) % R
    # C's discrete log from the verification equation itself
    sA, sB, sC = qd._scales(k, lcm)
    lag = _dl(native, native.fr_ap_lagrange_dev(k, x), k)
    pub_term = 0
    for i in keys.pub_idx:                                                # val_i = beta*A_i(x) + alpha*B_i(x) + C_i(x)
        ai = sum(row.get(i, 0) * lag[g] for g, row in enumerate(ra)) % R * sA
        bi = sum(row.get(i, 0) * lag[g] for g, row in enumerate(rb)) % R * sB
        ci = sum(row.get(i, 0) * lag[g] for g, row in enumerate(rc)) % R * sC
        pub_term += wit[i] * (beta * ai + alpha * bi + ci)
    c = (a * b - alpha * beta - pub_term) % R * pow(delta, -1, R) % R
    assert _pt(A) == bn254.g1_mul(bn254.G1, a)
    assert _pt2(B) == bn254.g2_mul(bn254.G2, b)
    assert _pt(C) == bn254.g1_mul(bn254.G1, c)
    # and sigma1_3 on the public wires is val_i / gamma
    i = keys.pub_idx[-1]
    ai = sum(row.get(i, 0) * lag[g] for g, row in enumerate(ra)) % R * sA
    bi = sum(row.get(i, 0) * lag[g] for g, row in enumerate(rb)) % R * sB
    ci = sum(row.get(i, 0) * lag[g] for g, row in enumerate(rc)) % R * sC
    assert _pt(keys.sigma1_3[i]) == bn254.g1_mul(bn254.G1, (beta * ai + alpha * bi + ci) * pow(gamma, -1, R) % R)


def test_small_circuit_verifies_with_pairings(native):
    """End to end with the verifier mirror (GPU public-input MSM + the oracle's pairing): accept, and reject
    a tampered public input."""
    from oracle import plonk_verifier
    from interactive_zkp_study_b200.zkp.groth16 import qap_device as qd, verifying
    bn = plonk_verifier._pairing()

    def oracle_pairing(q, p):
        q2 = (bn.FQ2([int(q[0].coeffs[0]), int(q[0].coeffs[1])]), bn.FQ2([int(q[1].coeffs[0]), int(q[1].coeffs[1])]))
        p1 = None if p is None else (bn.FQ(int(p[0])), bn.FQ(int(p[1])))
        return bn.pairing(q2, p1)
    k = 12
    ra, rb, rc, wit = _synthetic_circuit(k, 77)
    dev = qd.DeviceR1CS(qd.SparseR1CS.from_rows(k, k + 2, ra, rb, rc))
    keys = qd.setup(dev, 11, 22, 33, 44, 55, precompute=False)
    w = native.scalars_load(native.fr_vec_bytes(wit), k + 2)
    A, B, C = qd.prove(keys, dev, w, 6, 7)
    rx_pub = [(i, wit[i]) for i in keys.pub_idx]
    assert verifying.verify(A, B, C, keys.sigma1_1, keys.sigma1_3, keys.sigma2_1, rx_pub, pairing=oracle_pairing) is True
    bad = [(0, 1), (1, wit[1] + 1)]
    assert verifying.verify(A, B, C, keys.sigma1_1, keys.sigma1_3, keys.sigma2_1, bad, pairing=oracle_pairing) is False
