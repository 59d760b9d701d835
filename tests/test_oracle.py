"""CPU: the oracle (oracle/bn254.py, oracle/ref_path.py, oracle/shim) against the golden vectors
minted by the reference's own modules (tests/golden/make_golden.py) and public known answers.
This is what pins the oracle; the GPU suites then compare the CUDA path with the oracle."""
import hashlib

from oracle import bn254, ref_path, synthetic
from tests.util import g1, g2, ints, load

R = bn254.R


def test_public_known_answers():
    # EIP-196 test vector: 2 * G1
    assert bn254.g1_double(bn254.G1) == (
        1368015179489954701390400359078579693043519447331113978918064868415326638035,
        9918110051302171585080402603319702774565515993150576347155970296011118125764)
    assert bn254.g1_mul(bn254.G1, R) is None and bn254.g2_mul(bn254.G2, R) is None
    assert bn254.g1_is_on_curve(bn254.G1) and bn254.g2_is_on_curve(bn254.G2)
    assert (R - 1) % (1 << 28) == 0 and ((R - 1) >> 28) % 2 == 1
    # SURVEY 8c KATs: omega_4 and tau(seed=42)
    assert ref_path.get_root_of_unity(4) == 21888242871839275217838484774961031246007050428528088939761107053157389710902
    assert int.from_bytes(hashlib.sha256(b"42").digest(), "big") % R == \
        8365577799539384663899794442022354891237484320765090705979616311134436098119


def test_group_edge_semantics():
    P = bn254.g1_mul(bn254.G1, 77)
    assert bn254.g1_add(None, P) == P and bn254.g1_add(P, None) == P
    assert bn254.g1_add(P, P) == bn254.g1_double(P)
    assert bn254.g1_add(P, bn254.g1_neg(P)) is None
    assert bn254.g1_mul(P, 0) is None and bn254.g1_mul(P, 1) == P and bn254.g1_mul(None, 5) is None
    assert bn254.g1_mul(P, R + 3) == bn254.g1_mul(P, 3)
    Q = bn254.g2_mul(bn254.G2, 9)
    assert bn254.g2_add(Q, bn254.g2_neg(Q)) is None and bn254.g2_add(Q, Q) == bn254.g2_double(Q)


def test_fft_family_matches_reference():
    for case in load("primitives.json")["fft"]:
        v, w = ints(case["in"]), int(case["omega"])
        assert w == ref_path.get_root_of_unity(case["n"])
        assert ref_path.fft(v, w) == ints(case["fft"])
        assert ref_path.ifft(v, w) == ints(case["ifft"])
        assert ref_path.coset_fft(v, w) == ints(case["coset_fft"])
        assert ref_path.coset_ifft(v, w) == ints(case["coset_ifft"])
        assert ref_path.coset_fft(v, w, 7) == ints(case["coset_fft_k7"])


def test_poly_mul_div_match_reference():
    p = load("primitives.json")
    for case in p["poly_div"]:
        q, r = ref_path.poly_div(ints(case["a"]), ints(case["b"]))
        assert q == ints(case["q"]) and r == ints(case["r"])
    for case in p["poly_mul"]:
        assert ref_path.poly_mul(ints(case["a"]), ints(case["b"])) == ints(case["ab"])


def test_srs_and_commit_match_reference():
    p = load("primitives.json")
    g1p, g2p, tau = ref_path.srs_generate(12, 42)
    assert g1p == [g1(x) for x in p["srs42"]["g1_powers"]]
    assert g2p == [g2(x) for x in p["srs42"]["g2_powers"]]
    for case in p["commit"]:
        assert ref_path.commit(ints(case["coeffs"]), g1p) == g1(case["commitment"])


def test_groth16_toy_matches_reference():
    g = load("groth16_toy.json")
    Ax, Bx, Cx = ([ints(r) for r in g[k]] for k in ("Ax", "Bx", "Cx"))
    Zx, Rx = ints(g["Zx"]), ints(g["Rx"])
    Hx, rem = ref_path.hxr(Ax, Bx, Cx, Zx, Rx)
    assert Hx == ints(g["Hx"]) and rem == ints(g["remainder"])
    s11, s12 = [g1(p) for p in g["sigma1_1"]], [g1(p) for p in g["sigma1_2"]]
    s14, s15 = [g1(p) for p in g["sigma1_4"]], [g1(p) for p in g["sigma1_5"]]
    s21, s22 = [g2(p) for p in g["sigma2_1"]], [g2(p) for p in g["sigma2_2"]]
    r, s = int(g["r"]), int(g["s"])
    A = ref_path.proof_a(s11, s12, Ax, Rx, r)
    assert A == g1(g["proof_a"])
    assert ref_path.proof_b(s21, s22, Bx, Rx, s) == g2(g["proof_b"])
    assert ref_path.proof_c(s11, s12, s14, s15, Bx, Rx, Hx, s, r, A, g["pub_r_indexs"]) == g1(g["proof_c"])
    # the relation zkp/groth16/test.py:331 checks: proof_A == (alpha + A(x) + r*delta) * G1
    t = {k: int(v) for k, v in g["toxic"].items()}
    u = ref_path.g16_multiply_vec_matrix(Rx, Ax)
    a_at_x = ref_path.poly_eval(u[:g["numGates"]], t["x_val"])
    assert A == bn254.g1_mul(bn254.G1, (t["alpha"] + a_at_x + r * t["delta"]) % R)


def test_plonk_accumulator_and_openings_match_reference():
    for name in ("plonk_x3.json", "plonk_n1.json", "plonk_chain16.json"):
        f = load(name)
        n, w, dom = f["n"], int(f["omega"]), ints(f["domain"])
        assert dom == ref_path.get_roots_of_unity(n)
        s1, s2, s3 = (ref_path.fft(ints(f["pre"][k]) + [0] * (n - len(f["pre"][k])), w)
                      for k in ("s_sigma1", "s_sigma2", "s_sigma3"))
        beta, gamma = int(f["challenges"]["beta"]), int(f["challenges"]["gamma"])
        z = ref_path.compute_accumulator(ints(f["a_vals"]), ints(f["b_vals"]), ints(f["c_vals"]), s1, s2, s3, n, dom,
                                         beta, gamma)
        blinds = ints(f["blinds"])
        zh = [R - 1] + [0] * (n - 1) + [1]
        z_poly = ref_path.poly_add(ref_path.ifft(z, w), ref_path.poly_mul(ref_path.trim(blinds[6:9]), zh))
        assert z_poly == ints(f["polys"]["z"])
        g1p = [g1(p) for p in f["g1_powers"]]
        assert ref_path.commit(z_poly, g1p) == g1(f["proof"]["z_comm"])
        zeta = int(f["challenges"]["zeta"])
        assert ref_path.poly_eval(ints(f["polys"]["a"]), zeta) == int(f["proof"]["a_eval"])
        assert ref_path.poly_eval(z_poly, zeta * w % R) == int(f["proof"]["z_omega_eval"])


def test_synthetic_stream_is_reduced_and_deterministic():
    v = synthetic.scalars(0x5EED0001, 2000)
    assert all(0 <= x < R for x in v)
    assert v[:3] == [synthetic.scalar(0x5EED0001, i) for i in range(3)]
    assert v != synthetic.scalars(0x5EED0002, 2000)


def _golden_proof(f):
    proof = {k: (g1(v) if k.endswith("_comm") else int(v)) for k, v in f["proof"].items()}
    pre = {k: g1(v) for k, v in f["pre_comm"].items()}
    return proof, pre


def test_oracle_plonk_verifier_accepts_golden_and_rejects_tampering():
    """The restated verifier (oracle/plonk_verifier.py) on the reference-minted proofs: accepts each,
    rejects a tampered evaluation and a tampered commitment (reference tests/plonk/test_e2e.py:198-250)."""
    from oracle import plonk_verifier
    for name in ("plonk_n1.json", "plonk_x3.json", "plonk_chain16.json"):
        f = load(name)
        proof, pre = _golden_proof(f)
        g2p = [g2(p) for p in f["g2_powers"]]
        assert plonk_verifier.verify(proof, pre, f["n"], int(f["omega"]), g2p) is True
        if name == "plonk_x3.json":
            bad = dict(proof)
            bad["a_eval"] = (bad["a_eval"] + 1) % R
            assert plonk_verifier.verify(bad, pre, f["n"], int(f["omega"]), g2p) is False
            bad = dict(proof)
            bad["z_comm"] = bn254.g1_add(bad["z_comm"], bn254.G1)
            assert plonk_verifier.verify(bad, pre, f["n"], int(f["omega"]), g2p) is False


def test_qap_construction_matches_reference():
    """oracle r1cs_to_qap_times_lcm (exact Fr) == the reference's float-Lagrange x determinant pipeline
    (qap_creator_lcm.py:114-135 + getFRPoly2D) on the R1CS the reference's own compiler produced."""
    for c in load("groth16_qap.json")["cases"]:
        Ax, Bx, Cx, Z = ref_path.r1cs_to_qap_times_lcm(c["r1cs_A"], c["r1cs_B"], c["r1cs_C"])
        w, k, m = c["witness"], c["numGates"], c["numWires"]
        assert Z == ints(c["Zx"]) and len(Ax) == m and len(Ax[0]) == k
        for M, key in ((Ax, "uA"), (Bx, "uB"), (Cx, "uC")):
            assert ref_path.g16_multiply_vec_matrix(w, M)[:k] == ints(c[key])
        x = int(c["toxic"]["x_val"])
        assert ref_path.qap_eval_rows(Ax, x) == ints(c["Ax_val"])
        assert ref_path.qap_eval_rows(Bx, x) == ints(c["Bx_val"])
        assert ref_path.qap_eval_rows(Cx, x) == ints(c["Cx_val"])
        assert ref_path.poly_eval(Z, x) == int(c["Zx_val"])
        Hx, rem = ref_path.hxr(Ax, Bx, Cx, Z, w)
        assert Hx == ints(c["Hx"]) and not any(rem)
