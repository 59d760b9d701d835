"""Round 4 (/root/reference/zkp/plonk/prover/round4.py:39-81): challenge zeta and six openings,
each a two-level Horner evaluation on the GPU."""


def execute(state):
    state.zeta = state.transcript.challenge_scalar(b"zeta")
    zeta, pp, proof = state.zeta, state.preprocessed, state.proof
    proof.a_eval = state.a_poly.evaluate(zeta)
    proof.b_eval = state.b_poly.evaluate(zeta)
    proof.c_eval = state.c_poly.evaluate(zeta)
    proof.s_sigma1_eval = pp.s_sigma1_poly.evaluate(zeta)
    proof.s_sigma2_eval = pp.s_sigma2_poly.evaluate(zeta)
    proof.z_omega_eval = state.z_poly.evaluate(zeta * state.omega)
    for name in ("a_eval", "b_eval", "c_eval", "s_sigma1_eval", "s_sigma2_eval", "z_omega_eval"):
        state.transcript.append_scalar(name.encode(), getattr(proof, name))
