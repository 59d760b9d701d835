"""Round 1 (/root/reference/zkp/plonk/prover/round1.py:38-108): three iNTTs, blinding by
(b1 x + b0) * Z_H, three commitments -- iNTT, product and MSM on the GPU."""
import secrets

from ..field import FR, CURVE_ORDER
from ..polynomial import Polynomial
from ..kzg import commit


def _add_blinding(poly, zh, num_blinds):
    blind = Polynomial([FR(secrets.randbelow(CURVE_ORDER)) for _ in range(num_blinds)])
    return poly + blind * zh


def execute(state):
    zh = Polynomial.vanishing(state.n)
    state.pi_poly = Polynomial.zero()
    for wire in ("a", "b", "c"):
        poly = Polynomial.from_evaluations(getattr(state, wire + "_vals"), state.omega)
        setattr(state, wire + "_poly", _add_blinding(poly, zh, 2))
    for wire in ("a", "b", "c"):
        setattr(state.proof, wire + "_comm", commit(getattr(state, wire + "_poly"), state.srs))
    for wire in ("a", "b", "c"):
        state.transcript.append_point(wire.encode() + b"_comm", getattr(state.proof, wire + "_comm"))
