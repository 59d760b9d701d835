"""Drop-in for /root/reference/zkp/plonk/preprocessor.py:59-130: eight iNTTs (selector and
permutation polynomials) and eight KZG commitments, all on the GPU through Polynomial / commit.
Pads ``circuit.gates`` to a power of two exactly as the reference does (:85-88)."""
from .field import FR, get_root_of_unity, get_roots_of_unity
from .polynomial import Polynomial
from .kzg import commit
from .permutation import build_permutation_polynomials
from .utils import next_power_of_2


class PreprocessedData:
    pass


def _zero_gate(circuit):
    # the padding gate is a Gate(0,0,0,0,0) of the circuit's own Gate class (reference :86-88)
    gate_cls = type(circuit.gates[0]) if circuit.gates else None
    if gate_cls is None:
        raise ValueError("cannot pad an empty circuit")
    return gate_cls(FR(0), FR(0), FR(0), FR(0), FR(0))


def preprocess(circuit, srs):
    result = PreprocessedData()
    n = next_power_of_2(circuit.n)
    while len(circuit.gates) < n:
        circuit.gates.append(_zero_gate(circuit))
    result.n = n
    result.omega = get_root_of_unity(n)
    result.domain = get_roots_of_unity(n)
    names = ("q_l", "q_r", "q_o", "q_m", "q_c")
    for name, evals in zip(names, circuit.get_selector_polynomials()):
        poly = Polynomial.from_evaluations(evals, result.omega)
        setattr(result, name + "_poly", poly)
    for name in names:
        setattr(result, name + "_comm", commit(getattr(result, name + "_poly"), srs))
    result.sigma = circuit.build_copy_constraints()
    for name, evals in zip(("s_sigma1", "s_sigma2", "s_sigma3"),
                           build_permutation_polynomials(result.sigma, n, result.domain)):
        setattr(result, name + "_poly", Polynomial.from_evaluations(evals, result.omega))
    for name in ("s_sigma1", "s_sigma2", "s_sigma3"):
        setattr(result, name + "_comm", commit(getattr(result, name + "_poly"), srs))
    result.num_public_inputs = circuit.num_public_inputs
    return result
