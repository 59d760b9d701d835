"""-m gpu: verifier-side group work on the GPU: the Groth16 public-wire MSM and the
PLONK verifier's two linear combinations, with the pairings supplied by the oracle's py_ecc restatement.
Accept / reject decisions on the reference-minted proofs, as in the reference's
tests/groth16/test_verifying.py:36-62 and tests/plonk/test_e2e.py:133-250."""
import copy

import pytest

from oracle import bn254, plonk_verifier
from tests.util import g1, g2, ints, load

pytestmark = pytest.mark.gpu
R = bn254.R


@pytest.fixture(scope="module")
def pairing():
    bn = plonk_verifier._pairing()

    def fn(q, p):
        q2 = (bn.FQ2([int(q[0].coeffs[0]), int(q[0].coeffs[1])]), bn.FQ2([int(q[1].coeffs[0]), int(q[1].coeffs[1])]))
        p1 = None if p is None else (bn.FQ(int(p[0])), bn.FQ(int(p[1])))
        return bn.pairing(q2, p1)
    return fn


def test_groth16_verify_accepts_and_rejects(native, pairing):
    from interactive_zkp_study_b200.compat import FQ, FR, g1_from_ints, g2_from_ints
    from interactive_zkp_study_b200.zkp.groth16.verifying import verify
    from interactive_zkp_study_b200.zkp.groth16.proving import build_rpub_enum
    g = load("groth16_toy.json")
    A, B, C = g1_from_ints(g1(g["proof_a"])), g2_from_ints(g2(g["proof_b"])), g1_from_ints(g1(g["proof_c"]))
    s11 = [g1_from_ints(g1(p)) for p in g["sigma1_1"]]
    s13 = [g1_from_ints(g1(p)) for p in g["sigma1_3"]]
    s21 = [g2_from_ints(g2(p)) for p in g["sigma2_1"]]
    Rx = [FR(int(x)) for x in g["Rx"]]
    rx_pub = build_rpub_enum(g["pub_r_indexs"], Rx)
    assert verify(A, B, C, s11, s13, s21, rx_pub, pairing=pairing) is True
    badC = g1_from_ints(bn254.g1_add(g1(g["proof_c"]), bn254.G1))
    assert verify(A, B, badC, s11, s13, s21, rx_pub, pairing=pairing) is False
    bad_pub = [(rx_pub[0][0], rx_pub[0][1]), (rx_pub[1][0], rx_pub[1][1] + FR(1))]   # wrong public input
    assert verify(A, B, C, s11, s13, s21, bad_pub, pairing=pairing) is False
    with pytest.raises(NotImplementedError):
        verify(A, B, C, s11, s13, s21, rx_pub)        # no py_ecc in this image and no pairing given: loud


class _PP:
    pass


@pytest.mark.parametrize("name", ["plonk_n1.json", "plonk_x3.json", "plonk_chain16.json"])
def test_plonk_verify_matches_reference_decisions(native, pairing, name):
    from interactive_zkp_study_b200.compat import g1_from_ints, g2_from_ints
    from interactive_zkp_study_b200.zkp.plonk.field import FR
    from interactive_zkp_study_b200.zkp.plonk.prover import Proof
    from interactive_zkp_study_b200.zkp.plonk.srs import SRS
    from interactive_zkp_study_b200.zkp.plonk.verifier import verify
    f = load(name)
    proof = Proof()
    for k, v in f["proof"].items():
        setattr(proof, k, g1_from_ints(g1(v)) if k.endswith("_comm") else FR(int(v)))
    pp = _PP()
    pp.n, pp.omega = f["n"], FR(int(f["omega"]))
    for k, v in f["pre_comm"].items():
        setattr(pp, k + "_comm", g1_from_ints(g1(v)))
    srs = SRS([], [g2_from_ints(g2(p)) for p in f["g2_powers"]], f["srs_max_degree"])
    assert verify(proof, [], pp, srs, pairing=pairing) is True
    for field in ("a_eval", "r_eval", "z_omega_eval"):                 # single-field tampering (test_e2e.py:198-250)
        bad = copy.deepcopy(proof)
        setattr(bad, field, getattr(bad, field) + FR(1))
        assert verify(bad, [], pp, srs, pairing=pairing) is False, field
    bad = copy.deepcopy(proof)
    bad.t_mid_comm = g1_from_ints(bn254.g1_add(g1(f["proof"]["t_mid_comm"]), bn254.G1))
    assert verify(bad, [], pp, srs, pairing=pairing) is False
