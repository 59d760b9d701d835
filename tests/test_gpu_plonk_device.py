"""-m gpu: the device-resident PLONK prover (zkp/plonk/device_prover.py, BASELINE config 4 shape).
(1) With the golden runs' blinding scalars it reproduces the reference-minted proofs bit for bit
(n = 1, 4, 16).  (2) On synthetic circuits up to 2^14 gates its proofs are accepted by the oracle's
restatement of the reference verifier (itself pinned on the golden proofs) and a broken witness raises
the reference's ValueError."""
import os
import sys

import pytest

from oracle import bn254, plonk_verifier
from tests.util import g1, g2, ints, load

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))

pytestmark = pytest.mark.gpu
R = bn254.R


def _pt(p):
    return None if p is None else (int(p[0]), int(p[1]))


def _proof_dict(proof):
    out = {}
    for k, v in vars(proof).items():
        out[k] = _pt(v) if k.endswith("_comm") else int(v)
    return out


@pytest.mark.parametrize("name", ["plonk_n1.json", "plonk_x3.json", "plonk_chain16.json"])
def test_device_prover_reproduces_golden_proofs(native, name):
    from interactive_zkp_study_b200.zkp.plonk import device_prover as dp
    import plonk_synth
    f = load(name)
    n = f["n"]
    pts = [g1(p) for p in f["g1_powers"]]
    srs = native.g1_table_load(native.g1_vec_bytes(pts), len(pts))
    sel = f["selectors"]
    pad = lambda v: ints(v) + [0] * (n - len(v))
    sel_h = [native.scalars_load(native.fr_vec_bytes(pad(sel[k])), n) for k in ("q_l", "q_r", "q_o", "q_m", "q_c")]
    sig_h = [native.scalars_load(b, n) for b in plonk_synth.sigma_evals_bytes(f["sigma"], n, int(f["omega"]))]
    key = dp.preprocess(n, sel_h, sig_h, srs, len(pts))
    for k in dp.CIRCUIT_POLYS:
        assert key.comm[k] == g1(f["pre_comm"][k]), k
        got = native.fr_vec_from_bytes(native.scalars_download(key.coeffs[k], 0, n))
        want = ints(f["pre"][k])
        assert got == want + [0] * (n - len(want)), k
    wit = [native.scalars_load(native.fr_vec_bytes(ints(f[k + "_vals"])), n) for k in "abc"]
    proof, st = dp.prove(key, *wit, blinds=ints(f["blinds"]), keep=True)
    for k, v in f["challenges"].items():
        assert st["challenges"][k] == int(v), k
    for k, v in f["proof"].items():
        got = getattr(proof, k)
        assert (_pt(got) == g1(v)) if k.endswith("_comm") else (int(got) == int(v)), k
    for k, length in (("a", n + 2), ("z", n + 3)):
        want = ints(f["polys"][k])
        got = native.fr_vec_from_bytes(native.scalars_download(st[k], 0, length))
        assert got == want + [0] * (length - len(want)), k
    t = native.fr_vec_from_bytes(native.scalars_download(st["t"], 0, 3 * n + 6))
    want_t = ints(f["polys"]["t_lo"]) + [0] * (n - len(f["polys"]["t_lo"])) + ints(f["polys"]["t_mid"]) + \
        [0] * (n - len(f["polys"]["t_mid"])) + ints(f["polys"]["t_hi"])
    assert t == want_t + [0] * (3 * n + 6 - len(want_t))


@pytest.mark.parametrize("log_n", [3, 8, 14])
def test_device_prover_synthetic_verifies(native, log_n):
    from interactive_zkp_study_b200.zkp.plonk import device_prover as dp
    import plonk_synth
    n = 1 << log_n
    circ = plonk_synth.chain_circuit(n, seed=log_n)
    key, wit, tau = plonk_synth.device_setup(circ)
    proof = dp.prove(key, *wit)
    pre = {k: key.comm[k] for k in dp.CIRCUIT_POLYS}
    g2p = [bn254.G2, bn254.g2_mul(bn254.G2, tau)]
    pd = _proof_dict(proof)
    assert all(bn254.g1_is_on_curve(v) for k, v in pd.items() if k.endswith("_comm"))
    assert plonk_verifier.verify(pd, pre, n, key.omega, g2p) is True
    bad = dict(pd)
    bad["c_eval"] = (bad["c_eval"] + 1) % R
    assert plonk_verifier.verify(bad, pre, n, key.omega, g2p) is False
    proof2 = dp.prove(key, *wit)                      # fresh blinding -> different proof, still valid
    assert _pt(proof2.a_comm) != _pt(proof.a_comm)


def test_device_prover_2_20_config4(native):
    """BASELINE config 4 inside the GPU suite: PLONK prove at 2^20 gates, accepted by the oracle's restatement of
    the reference verifier; a tampered evaluation is rejected."""
    from interactive_zkp_study_b200.zkp.plonk import device_prover as dp
    import plonk_synth
    n = 1 << 20
    key, wit, tau = plonk_synth.device_setup(plonk_synth.chain_circuit(n, seed=20))
    pd = _proof_dict(dp.prove(key, *wit))
    pre = {k: key.comm[k] for k in dp.CIRCUIT_POLYS}
    g2p = [bn254.G2, bn254.g2_mul(bn254.G2, tau)]
    assert plonk_verifier.verify(pd, pre, n, key.omega, g2p) is True
    bad = dict(pd)
    bad["z_omega_eval"] = (bad["z_omega_eval"] + 1) % R
    assert plonk_verifier.verify(bad, pre, n, key.omega, g2p) is False


def test_device_prover_rejects_bad_witness(native):
    from interactive_zkp_study_b200.zkp.plonk import device_prover as dp
    import plonk_synth
    circ = plonk_synth.chain_circuit(64, seed=9, break_witness=True)
    key, wit, _ = plonk_synth.device_setup(circ)
    with pytest.raises(ValueError, match="나누어 떨어지지 않"):
        dp.prove(key, *wit)
