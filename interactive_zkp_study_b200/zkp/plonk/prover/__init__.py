"""Drop-in for /root/reference/zkp/plonk/prover/__init__.py: Proof, ProverState, prove().
Field names are the reference's (:42-155) because the Flask routes and serializers read them."""
from ..transcript import Transcript
from . import round1, round2, round3, round4, round5

_PROOF_FIELDS = ("a_comm", "b_comm", "c_comm", "z_comm", "t_lo_comm", "t_mid_comm", "t_hi_comm",
                 "a_eval", "b_eval", "c_eval", "s_sigma1_eval", "s_sigma2_eval", "z_omega_eval",
                 "r_eval", "W_zeta_comm", "W_zeta_omega_comm")


class Proof:
    def __init__(self):
        for name in _PROOF_FIELDS:
            setattr(self, name, None)


class ProverState:
    def __init__(self, a_vals, b_vals, c_vals, public_inputs, preprocessed, srs):
        self.a_vals, self.b_vals, self.c_vals = a_vals, b_vals, c_vals
        self.public_inputs = public_inputs
        self.preprocessed = preprocessed
        self.srs = srs
        self.transcript = Transcript()
        self.n = preprocessed.n
        self.omega = preprocessed.omega
        self.domain = preprocessed.domain
        for name in ("a_poly", "b_poly", "c_poly", "z_poly", "t_lo_poly", "t_mid_poly", "t_hi_poly",
                     "beta", "gamma", "alpha", "zeta", "v", "pi_poly"):
            setattr(self, name, None)
        self.proof = Proof()

    def build_proof(self):
        return self.proof


def prove(circuit, a_vals, b_vals, c_vals, public_inputs, preprocessed, srs):
    state = ProverState(a_vals, b_vals, c_vals, public_inputs, preprocessed, srs)
    for rnd in (round1, round2, round3, round4, round5):
        rnd.execute(state)
    return state.build_proof()
