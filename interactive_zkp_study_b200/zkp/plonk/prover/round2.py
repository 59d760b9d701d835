"""Round 2 (/root/reference/zkp/plonk/prover/round2.py:45-85): challenges beta, gamma; the
grand-product accumulator (batch inversion + prefix product on the GPU), one iNTT, three blinding
scalars, one commitment."""
import secrets

from ..field import FR, CURVE_ORDER
from ..polynomial import Polynomial
from ..kzg import commit
from ..permutation import compute_accumulator


def execute(state):
    state.beta = state.transcript.challenge_scalar(b"beta")
    state.gamma = state.transcript.challenge_scalar(b"gamma")
    z_evals = compute_accumulator(state.a_vals, state.b_vals, state.c_vals, state.preprocessed.sigma, state.n,
                                  state.domain, state.beta, state.gamma)
    z_poly = Polynomial.from_evaluations(z_evals, state.omega)
    blind = Polynomial([FR(secrets.randbelow(CURVE_ORDER)) for _ in range(3)])
    state.z_poly = z_poly + blind * Polynomial.vanishing(state.n)
    state.proof.z_comm = commit(state.z_poly, state.srs)
    state.transcript.append_point(b"z_comm", state.proof.z_comm)
