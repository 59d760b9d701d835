"""-m gpu: the Groth16 drop-in functions (same signatures as /root/reference/zkp/groth16/proving.py
and poly_utils.py) on the reference's toy circuit: bit-exact with the golden values the reference
itself produced (tests/golden/groth16_toy.json; accepted there by the reference's verifier), modelled
on /root/reference/tests/groth16/test_proving.py, test_poly_utils.py:136-148, test_integration.py."""
import pytest

from oracle import bn254, ref_path
from tests.util import g1, g2, ints, load

pytestmark = pytest.mark.gpu
R = bn254.R


@pytest.fixture(scope="module")
def toy(native):
    from interactive_zkp_study_b200.compat import FQ, FQ2, FR
    g = load("groth16_toy.json")
    pt1 = lambda p: (FQ(int(p[0])), FQ(int(p[1])))
    pt2 = lambda p: (FQ2([int(p[0][0]), int(p[0][1])]), FQ2([int(p[1][0]), int(p[1][1])]))
    d = {"g": g}
    for k in ("Ax", "Bx", "Cx"):
        d[k] = [[FR(int(x)) for x in row] for row in g[k]]
    d["Zx"] = [FR(int(x)) for x in g["Zx"]]
    d["Rx"] = [FR(int(x)) for x in g["Rx"]]
    for k in ("sigma1_1", "sigma1_2", "sigma1_4", "sigma1_5"):
        d[k] = [pt1(p) for p in g[k]]
    for k in ("sigma2_1", "sigma2_2"):
        d[k] = [pt2(p) for p in g[k]]
    return d


def _g1i(p):
    return None if p is None else (int(p[0]), int(p[1]))


def _g2i(p):
    return None if p is None else ((int(p[0].coeffs[0]), int(p[0].coeffs[1])), (int(p[1].coeffs[0]), int(p[1].coeffs[1])))


def test_hxr_matches_reference(toy):
    from interactive_zkp_study_b200.zkp.groth16 import poly_utils as pu
    from interactive_zkp_study_b200.compat import FR
    g = toy["g"]
    Hx, rem = pu.hxr(toy["Ax"], toy["Bx"], toy["Cx"], toy["Zx"], g["R_raw"])
    assert [int(x) for x in Hx] == ints(g["Hx"])                  # [-528, 2456, -496, 0, 0, 0, 0] mod r
    assert len(Hx) == 2 * g["numWires"] - g["numGates"] - 1        # the reference's length quirk
    assert [int(x) for x in rem] == ints(g["remainder"]) == [0] * g["numGates"]
    assert all(isinstance(x, FR) for x in Hx)
    # polynomial identity at x_val (reference tests/groth16/test_integration.py:44-56)
    x = int(g["toxic"]["x_val"])
    ra, rb, rc = (pu._multiply_vec_matrix(toy["Rx"], toy[k]) for k in ("Ax", "Bx", "Cx"))
    lhs = (int(pu._eval_poly(ra, x)) * int(pu._eval_poly(rb, x)) - int(pu._eval_poly(rc, x))) % R
    assert lhs == int(pu.hx_val(Hx, x)) * int(pu.zx_val(toy["Zx"], x)) % R


def test_poly_helpers_match_reference_semantics(toy):
    from interactive_zkp_study_b200.zkp.groth16 import poly_utils as pu
    a, b = [3, 0, R - 1, 7], [5, 9]
    assert [int(x) for x in pu._multiply_polys(a, b)] == ref_path.g16_multiply_polys(a, b)
    assert [int(x) for x in pu._subtract_polys(a, b)] == ref_path.g16_subtract_polys(a, b)
    q, r = pu._div_polys(a, b)
    wq, wr = ref_path.g16_div_polys(a, b)
    assert [int(x) for x in q] == wq and [int(x) for x in r] == wr
    with pytest.raises(AssertionError):
        pu._multiply_vec_matrix([1, 2], [[1, 2], [3, 4]])          # reference assert, poly_utils.py:54


def test_proofs_bit_exact_with_reference(toy):
    from interactive_zkp_study_b200.zkp.groth16.proving import proof_a, proof_b, proof_c, build_rpub_enum
    from interactive_zkp_study_b200.compat import FQ, FR
    g = toy["g"]
    r, s = FR(int(g["r"])), FR(int(g["s"]))
    A = proof_a(toy["sigma1_1"], toy["sigma1_2"], toy["Ax"], toy["Rx"], r)
    assert isinstance(A, tuple) and len(A) == 2 and isinstance(A[0], FQ)
    assert _g1i(A) == g1(g["proof_a"])
    B = proof_b(toy["sigma2_1"], toy["sigma2_2"], toy["Bx"], toy["Rx"], s)
    assert _g2i(B) == g2(g["proof_b"])
    Hx = [FR(int(x)) for x in g["Hx"]]
    C = proof_c(toy["sigma1_1"], toy["sigma1_2"], toy["sigma1_4"], toy["sigma1_5"], toy["Bx"], toy["Rx"], Hx, s, r, A,
                pub_r_indexs=g["pub_r_indexs"])
    assert _g1i(C) == g1(g["proof_c"])
    # default pub_r_indexs == [0, 1] (proving.py:49-50)
    C2 = proof_c(toy["sigma1_1"], toy["sigma1_2"], toy["sigma1_4"], toy["sigma1_5"], toy["Bx"], toy["Rx"], Hx, s, r, A)
    assert _g1i(C2) == g1(g["proof_c"])
    assert bn254.g1_is_on_curve(_g1i(A)) and bn254.g2_is_on_curve(_g2i(B)) and bn254.g1_is_on_curve(_g1i(C))
    assert build_rpub_enum([0, 1], toy["Rx"]) == [(0, toy["Rx"][0]), (1, toy["Rx"][1])]


def test_proofs_match_oracle_for_other_randomness(toy):
    """Different r, s (raw ints larger than the curve order, as the Flask form can pass them,
    app.py:1253-1254) against the oracle's restatement of the nested loops."""
    from interactive_zkp_study_b200.zkp.groth16.proving import proof_a, proof_b, proof_c
    g = toy["g"]
    r, s = R + 12345, (1 << 255) + 99
    Ai = [ints(row) for row in g["Ax"]]
    Bi = [ints(row) for row in g["Bx"]]
    Rx = ints(g["Rx"])
    s11, s12 = [g1(p) for p in g["sigma1_1"]], [g1(p) for p in g["sigma1_2"]]
    s14, s15 = [g1(p) for p in g["sigma1_4"]], [g1(p) for p in g["sigma1_5"]]
    s21, s22 = [g2(p) for p in g["sigma2_1"]], [g2(p) for p in g["sigma2_2"]]
    A = proof_a(toy["sigma1_1"], toy["sigma1_2"], toy["Ax"], toy["Rx"], r)
    wantA = ref_path.proof_a(s11, s12, Ai, Rx, r)
    assert _g1i(A) == wantA
    assert _g2i(proof_b(toy["sigma2_1"], toy["sigma2_2"], toy["Bx"], toy["Rx"], s)) == ref_path.proof_b(s21, s22, Bi, Rx, s)
    Hx = ints(g["Hx"])
    C = proof_c(toy["sigma1_1"], toy["sigma1_2"], toy["sigma1_4"], toy["sigma1_5"], toy["Bx"], toy["Rx"], Hx, s, r, A)
    assert _g1i(C) == ref_path.proof_c(s11, s12, s14, s15, Bi, Rx, Hx, s, r, wantA)


def test_setup_sigmas_bit_exact_with_reference(toy):
    """CRS generation (reference zkp/groth16/setup.py:15-69; tests/groth16/test_setup.py) with the
    conftest toxic waste: every sigma list equals the reference's, placeholders included."""
    from interactive_zkp_study_b200.zkp.groth16 import setup as st
    from interactive_zkp_study_b200.zkp.groth16 import poly_utils as pu
    from interactive_zkp_study_b200.compat import FQ, FR
    g = toy["g"]
    t = {k: FR(int(v)) for k, v in g["toxic"].items()}
    k, m, pub = g["numGates"], g["numWires"], g["pub_r_indexs"]
    Axv, Bxv, Cxv = pu.ax_val(toy["Ax"], t["x_val"]), pu.bx_val(toy["Bx"], t["x_val"]), pu.cx_val(toy["Cx"], t["x_val"])
    Zxv = pu.zx_val(toy["Zx"], t["x_val"])
    assert [_g1i(p) for p in st.sigma11(t["alpha"], t["beta"], t["delta"])] == [g1(p) for p in g["sigma1_1"]]
    assert [_g1i(p) for p in st.sigma12(k, t["x_val"])] == [g1(p) for p in g["sigma1_2"]]
    s13, VAL = st.sigma13(m, t["alpha"], t["beta"], t["gamma"], Axv, Bxv, Cxv, pub_r_indexs=pub)
    assert [_g1i(p) for p in s13] == [g1(p) for p in g["sigma1_3"]]
    assert [int(v) for v in VAL] == ints(g["VAL"])
    s14 = st.sigma14(m, t["alpha"], t["beta"], t["delta"], Axv, Bxv, Cxv, pub_r_indexs=pub)
    assert [_g1i(p) for p in s14] == [g1(p) for p in g["sigma1_4"]]
    assert _g1i(s14[0]) == (0, 0) and isinstance(s14[0][0], FQ)        # placeholder, setup.py:50
    assert [_g1i(p) for p in st.sigma15(k, t["delta"], t["x_val"], Zxv)] == [g1(p) for p in g["sigma1_5"]]
    assert [_g2i(p) for p in st.sigma21(t["beta"], t["delta"], t["gamma"])] == [g2(p) for p in g["sigma2_1"]]
    assert [_g2i(p) for p in st.sigma22(k, t["x_val"])] == [g2(p) for p in g["sigma2_2"]]
