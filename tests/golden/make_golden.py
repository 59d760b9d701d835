#!/usr/bin/env python3
"""Generates the golden fixtures under tests/golden/ by running the REFERENCE's own modules
(imported from /root/reference) with oracle/shim standing in for the absent py-ecc==7.0.1.

Only runs in the build container (the GPU box has no /root/reference); the JSON it writes is
committed.  Integers are decimal strings; G1 points are [x, y], G2 points [[x0, x1], [y0, y1]],
infinity is null.

    python tests/golden/make_golden.py
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "shim"))
sys.path.insert(0, "/root/reference")
sys.setrecursionlimit(10000)


def s(x):
    return str(int(x))


def svec(v):
    return [s(x) for x in v]


def g1(p):
    return None if p is None else [s(p[0]), s(p[1])]


def g2(p):
    if p is None:
        return None
    return [[s(p[0].coeffs[0]), s(p[0].coeffs[1])], [s(p[1].coeffs[0]), s(p[1].coeffs[1])]]


def dump(name, obj):
    path = os.path.join(HERE, name)
    with open(path, "w") as fh:
        json.dump(obj, fh, indent=0, separators=(",", ":"))
    print("wrote", name, os.path.getsize(path), "bytes")


# ------------------------------------------------------------------ Groth16 toy pipeline
def groth16_toy():
    """Same pipeline and constants as /root/reference/tests/groth16/conftest.py:39-158."""
    from zkp.groth16.code_to_r1cs import code_to_r1cs_with_inputs, initialize_symbol
    from zkp.groth16.qap_creator_lcm import r1cs_to_qap_times_lcm
    from zkp.groth16 import poly_utils as pu
    from zkp.groth16 import setup as st
    from zkp.groth16.proving import proof_a, proof_b, proof_c, build_rpub_enum, FR
    from zkp.groth16.verifying import verify

    code = "\ndef qeval(x):\n    y = x**3\n    return y + x + 5\n"
    initialize_symbol()
    r, A, B, C = code_to_r1cs_with_inputs(code, [3])
    Ap, Bp, Cp, Z = r1cs_to_qap_times_lcm(A, B, C)
    alpha, beta, gamma, delta, x_val = FR(3926), FR(3604), FR(2971), FR(1357), FR(3721)
    Ax, Bx, Cx = pu.getFRPoly2D(Ap), pu.getFRPoly2D(Bp), pu.getFRPoly2D(Cp)
    Zx, Rx = pu.getFRPoly1D(Z), pu.getFRPoly1D(r)
    Hx, rem = pu.hxr(Ax, Bx, Cx, Zx, r)
    k, m = pu.getNumGates(Ax), pu.getNumWires(Ax)
    Axv, Bxv, Cxv, Zxv = pu.ax_val(Ax, x_val), pu.bx_val(Bx, x_val), pu.cx_val(Cx, x_val), pu.zx_val(Zx, x_val)
    pub = [0, 1]
    s11 = st.sigma11(alpha, beta, delta)
    s12 = st.sigma12(k, x_val)
    s13, VAL = st.sigma13(m, alpha, beta, gamma, Axv, Bxv, Cxv, pub_r_indexs=pub)
    s14 = st.sigma14(m, alpha, beta, delta, Axv, Bxv, Cxv, pub_r_indexs=pub)
    s15 = st.sigma15(k, delta, x_val, Zxv)
    s21 = st.sigma21(beta, delta, gamma)
    s22 = st.sigma22(k, x_val)
    rp, sp = FR(4106), FR(4565)
    A_ = proof_a(s11, s12, Ax, Rx, rp)
    B_ = proof_b(s21, s22, Bx, Rx, sp)
    C_ = proof_c(s11, s12, s14, s15, Bx, Rx, Hx, sp, rp, A_, pub_r_indexs=pub)
    ok = verify(A_, B_, C_, s11, s13, s21, build_rpub_enum(pub, Rx))
    assert ok is True
    # the valid numeric comment of the reference: VAL[0], VAL[5] of backend.py:363-367
    dump("groth16_toy.json", {
        "source": "reference modules zkp.groth16.* on oracle/shim; constants of tests/groth16/conftest.py:39-56",
        "toxic": {"alpha": "3926", "beta": "3604", "gamma": "2971", "delta": "1357", "x_val": "3721"},
        "r": "4106", "s": "4565", "pub_r_indexs": pub, "numGates": k, "numWires": m,
        "Ax": [svec(row) for row in Ax], "Bx": [svec(row) for row in Bx], "Cx": [svec(row) for row in Cx],
        "Zx": svec(Zx), "Rx": svec(Rx), "R_raw": [int(v) for v in r],
        "Hx": svec(Hx), "remainder": svec(rem),
        "sigma1_1": [g1(p) for p in s11], "sigma1_2": [g1(p) for p in s12], "sigma1_3": [g1(p) for p in s13],
        "sigma1_4": [g1(p) for p in s14], "sigma1_5": [g1(p) for p in s15],
        "sigma2_1": [g2(p) for p in s21], "sigma2_2": [g2(p) for p in s22],
        "VAL": svec(VAL),
        "proof_a": g1(A_), "proof_b": g2(B_), "proof_c": g1(C_), "verify": ok,
    })


# ------------------------------------------------------------------ Groth16 from the R1CS (QAP construction)
def groth16_qap():
    """R1CS -> QAP -> CRS -> proof for two programs, keeping the R1CS itself so that the sparse /
    interpolating device path (zkp/groth16/qap_device.py) can be checked against every stage of the
    reference's float-Lagrange + determinant pipeline (qap_creator_lcm.py:114-135)."""
    from zkp.groth16.code_to_r1cs import code_to_r1cs_with_inputs, initialize_symbol
    from zkp.groth16.qap_creator_lcm import r1cs_to_qap_times_lcm
    from zkp.groth16 import poly_utils as pu
    from zkp.groth16 import setup as st
    from zkp.groth16.proving import proof_a, proof_b, proof_c, build_rpub_enum, FR
    from zkp.groth16.verifying import verify
    cases = []
    programs = [
        ("x3_plus_x_plus_5", "\ndef qeval(x):\n    y = x**3\n    return y + x + 5\n", [3]),
        # six gates: the largest size at which the reference's float interpolation times det^2 stays exact
        # (at 8 gates its own remainder is already non-zero)
        ("six_gates", "\ndef qeval(x):\n    y = x**3\n    z = y * x + x\n    return z + y + 7\n", [2]),
    ]
    for idx, (name, code, inputs) in enumerate(programs):
        initialize_symbol()
        r, A, B, C = code_to_r1cs_with_inputs(code, inputs)
        Ap, Bp, Cp, Z = r1cs_to_qap_times_lcm(A, B, C)
        Ax, Bx, Cx = pu.getFRPoly2D(Ap), pu.getFRPoly2D(Bp), pu.getFRPoly2D(Cp)
        Zx, Rx = pu.getFRPoly1D(Z), pu.getFRPoly1D(r)
        k, m = pu.getNumGates(Ax), pu.getNumWires(Ax)
        alpha, beta, gamma, delta, x_val = (FR(v + 17 * idx) for v in (3926, 3604, 2971, 1357, 3721))
        Hx, rem = pu.hxr(Ax, Bx, Cx, Zx, r)
        uA = [sum((Rx[i] * Ax[i][j] for i in range(m)), FR(0)) for j in range(k)]
        uB = [sum((Rx[i] * Bx[i][j] for i in range(m)), FR(0)) for j in range(k)]
        uC = [sum((Rx[i] * Cx[i][j] for i in range(m)), FR(0)) for j in range(k)]
        Axv, Bxv, Cxv, Zxv = pu.ax_val(Ax, x_val), pu.bx_val(Bx, x_val), pu.cx_val(Cx, x_val), pu.zx_val(Zx, x_val)
        pub = [0, 1]
        s11 = st.sigma11(alpha, beta, delta)
        s12 = st.sigma12(k, x_val)
        s13, VAL = st.sigma13(m, alpha, beta, gamma, Axv, Bxv, Cxv, pub_r_indexs=pub)
        s14 = st.sigma14(m, alpha, beta, delta, Axv, Bxv, Cxv, pub_r_indexs=pub)
        s15 = st.sigma15(k, delta, x_val, Zxv)
        s21 = st.sigma21(beta, delta, gamma)
        s22 = st.sigma22(k, x_val)
        rp, sp = FR(4106 + idx), FR(4565 + idx)
        A_ = proof_a(s11, s12, Ax, Rx, rp)
        B_ = proof_b(s21, s22, Bx, Rx, sp)
        C_ = proof_c(s11, s12, s14, s15, Bx, Rx, Hx, sp, rp, A_, pub_r_indexs=pub)
        ok = verify(A_, B_, C_, s11, s13, s21, build_rpub_enum(pub, Rx))
        assert ok is True and all(int(v) == 0 for v in rem)
        cases.append({
            "name": name, "numGates": k, "numWires": m, "pub_r_indexs": pub,
            "r1cs_A": [[int(v) for v in row] for row in A], "r1cs_B": [[int(v) for v in row] for row in B],
            "r1cs_C": [[int(v) for v in row] for row in C], "witness": [int(v) for v in r],
            "toxic": {"alpha": s(alpha), "beta": s(beta), "gamma": s(gamma), "delta": s(delta), "x_val": s(x_val)},
            "r": s(rp), "s": s(sp),
            "uA": svec(uA), "uB": svec(uB), "uC": svec(uC), "Zx": svec(Zx), "Hx": svec(Hx),
            "Ax_val": svec(Axv), "Bx_val": svec(Bxv), "Cx_val": svec(Cxv), "Zx_val": s(Zxv),
            "sigma1_1": [g1(p) for p in s11], "sigma1_2": [g1(p) for p in s12], "sigma1_3": [g1(p) for p in s13],
            "sigma1_4": [g1(p) for p in s14], "sigma1_5": [g1(p) for p in s15],
            "sigma2_1": [g2(p) for p in s21], "sigma2_2": [g2(p) for p in s22],
            "proof_a": g1(A_), "proof_b": g2(B_), "proof_c": g1(C_), "verify": ok,
        })
    dump("groth16_qap.json", {
        "source": "reference zkp.groth16.{code_to_r1cs,qap_creator_lcm,poly_utils,setup,proving,verifying} on oracle/shim",
        "cases": cases})


# ------------------------------------------------------------------ PLONK
class SeededSecrets:
    """Stands in for the `secrets` module inside round1/round2 (they call secrets.randbelow)."""

    def __init__(self, seed):
        self.rng = random.Random(seed)
        self.drawn = []

    def randbelow(self, n):
        v = self.rng.randrange(n)
        self.drawn.append(v)
        return v


def chain_circuit(n_gates, seed):
    """A satisfiable circuit of alternating mul/add gates wired as a chain (uses only the reference's
    public Circuit API, /root/reference/zkp/plonk/circuit.py:124-205)."""
    from zkp.plonk.circuit import Circuit
    from zkp.plonk.field import FR
    rng = random.Random(seed)
    c = Circuit()
    a_vals, b_vals, c_vals = [], [], []
    prev = FR(rng.randrange(1, 1 << 40))
    for g in range(n_gates):
        other = FR(rng.randrange(1, 1 << 40))
        if g % 2 == 0:
            c.add_multiplication_gate()
            out = prev * other
        else:
            c.add_addition_gate()
            out = prev + other
        a_vals.append(prev)
        b_vals.append(other)
        c_vals.append(out)
        if g > 0:
            c.add_copy_constraint(g - 1, 2, g, 0)  # c_{g-1} == a_g
        prev = out
    return c, a_vals, b_vals, c_vals, []


def plonk_case(name, circuit, a_vals, b_vals, c_vals, public_inputs, srs_seed, blind_seed, tamper=False):
    from zkp.plonk.srs import SRS
    from zkp.plonk.preprocessor import preprocess
    from zkp.plonk.prover import ProverState, round1, round2, round3, round4, round5
    from zkp.plonk.verifier import verify
    from zkp.plonk.field import FR

    n_raw = circuit.n
    srs = SRS.generate(3 * max(n_raw, 1) + 10 + 8, seed=srs_seed)
    pp = preprocess(circuit, srs)
    n = pp.n
    # pad witness the way the reference's callers do (zeros for padding gates)
    a = list(a_vals) + [FR(0)] * (n - len(a_vals))
    b = list(b_vals) + [FR(0)] * (n - len(b_vals))
    c = list(c_vals) + [FR(0)] * (n - len(c_vals))
    sec = SeededSecrets(blind_seed)
    round1.secrets = sec
    round2.secrets = sec
    state = ProverState(a, b, c, public_inputs, pp, srs)
    for rnd in (round1, round2, round3, round4, round5):
        rnd.execute(state)
    proof = state.build_proof()
    ok = verify(proof, public_inputs, pp, srs)
    assert ok is True, name
    q_l, q_r, q_o, q_m, q_c = circuit.get_selector_polynomials()
    pf = proof
    out = {
        "source": "reference modules zkp.plonk.* on oracle/shim; secrets.randbelow replaced by random.Random(%d)" % blind_seed,
        "n": n, "n_raw": n_raw, "srs_seed": srs_seed, "srs_max_degree": srs.max_degree,
        "g1_powers": [g1(p) for p in srs.g1_powers], "g2_powers": [g2(p) for p in srs.g2_powers],
        "selectors": {"q_l": svec(q_l), "q_r": svec(q_r), "q_o": svec(q_o), "q_m": svec(q_m), "q_c": svec(q_c)},
        "sigma": [int(v) for v in pp.sigma], "num_public_inputs": pp.num_public_inputs,
        "omega": s(pp.omega), "domain": svec(pp.domain),
        "a_vals": svec(a), "b_vals": svec(b), "c_vals": svec(c), "public_inputs": svec(public_inputs),
        "blinds": svec(sec.drawn),
        "pre": {k: svec(getattr(pp, k + "_poly").coeffs) for k in
                ("q_l", "q_r", "q_o", "q_m", "q_c", "s_sigma1", "s_sigma2", "s_sigma3")},
        "pre_comm": {k: g1(getattr(pp, k + "_comm")) for k in
                     ("q_l", "q_r", "q_o", "q_m", "q_c", "s_sigma1", "s_sigma2", "s_sigma3")},
        "challenges": {k: s(getattr(state, k)) for k in ("beta", "gamma", "alpha", "zeta", "v")},
        "polys": {k: svec(getattr(state, k + "_poly").coeffs) for k in ("a", "b", "c", "z", "t_lo", "t_mid", "t_hi")},
        "proof": {
            **{k: g1(getattr(pf, k)) for k in ("a_comm", "b_comm", "c_comm", "z_comm", "t_lo_comm", "t_mid_comm",
                                               "t_hi_comm", "W_zeta_comm", "W_zeta_omega_comm")},
            **{k: s(getattr(pf, k)) for k in ("a_eval", "b_eval", "c_eval", "s_sigma1_eval", "s_sigma2_eval",
                                              "z_omega_eval", "r_eval")},
        },
        "verify": ok,
    }
    dump(name, out)


def plonk_all():
    from zkp.plonk.circuit import Circuit
    circuit, a, b, c, pub = Circuit.x3_plus_x_plus_5_eq_35()
    plonk_case("plonk_x3.json", circuit, a, b, c, pub, srs_seed=42, blind_seed=1001)
    # single multiplication gate (n = 1), as in /root/reference/tests/plonk/test_e2e.py:54-83
    from zkp.plonk.field import FR
    c1 = Circuit()
    c1.add_multiplication_gate()
    plonk_case("plonk_n1.json", c1, [FR(3)], [FR(4)], [FR(12)], [], srs_seed=7, blind_seed=1002)
    cc, a, b, c, pub = chain_circuit(13, seed=5)   # padded to n = 16
    plonk_case("plonk_chain16.json", cc, a, b, c, pub, srs_seed=12345, blind_seed=1003)


# ------------------------------------------------------------------ primitives
def primitives():
    from zkp.plonk.field import FR, get_root_of_unity, ec_mul, G1
    from zkp.plonk.polynomial import Polynomial, fft, ifft, poly_div
    from zkp.plonk.utils import coset_fft, coset_ifft
    from zkp.plonk.kzg import commit
    from zkp.plonk.srs import SRS
    from zkp.plonk.permutation import compute_accumulator
    rng = random.Random(2024)
    R = FR.field_modulus
    out = {"source": "reference zkp.plonk.{polynomial,utils,kzg,srs,permutation} on oracle/shim", "fft": [],
           "poly_div": [], "poly_mul": []}
    for n in (1, 2, 4, 8, 32):
        w = get_root_of_unity(n)
        v = [FR(rng.randrange(R)) for _ in range(n)]
        out["fft"].append({"n": n, "omega": s(w), "in": svec(v), "fft": svec(fft(v, w)), "ifft": svec(ifft(v, w)),
                           "coset_fft": svec(coset_fft(v, w)), "coset_ifft": svec(coset_ifft(v, w)),
                           "coset_fft_k7": svec(coset_fft(v, w, FR(7)))})
    for la, lb in ((5, 2), (9, 9), (12, 5), (3, 7)):
        a = Polynomial([FR(rng.randrange(R)) for _ in range(la)])
        b = Polynomial([FR(rng.randrange(R)) for _ in range(lb)])
        q, r = poly_div(a, b)
        out["poly_div"].append({"a": svec(a.coeffs), "b": svec(b.coeffs), "q": svec(q.coeffs), "r": svec(r.coeffs)})
        out["poly_mul"].append({"a": svec(a.coeffs), "b": svec(b.coeffs), "ab": svec((a * b).coeffs)})
    srs = SRS.generate(12, seed=42)
    polys = [[FR(rng.randrange(R)) for _ in range(k)] for k in (1, 4, 13)] + [[FR(0)], [FR(5), FR(0), FR(0), FR(9)]]
    out["srs42"] = {"max_degree": 12, "g1_powers": [g1(p) for p in srs.g1_powers], "g2_powers": [g2(p) for p in srs.g2_powers]}
    out["commit"] = [{"coeffs": svec(Polynomial(p).coeffs), "commitment": g1(commit(Polynomial(p), srs))} for p in polys]
    out["kat"] = {"omega4": s(get_root_of_unity(4)), "two_G1": g1(ec_mul(G1, 2))}
    dump("primitives.json", out)


if __name__ == "__main__":
    only = sys.argv[1:]
    for fn in (groth16_toy, groth16_qap, primitives, plonk_all):
        if not only or fn.__name__ in only:
            fn()
