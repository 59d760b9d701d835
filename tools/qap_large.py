"""Groth16 from a sparse R1CS at 2^k constraints: synthetic satisfiable circuit, CRS from
known toxic values, witness polynomials by sparse mat-vec + interpolation on {1..k}, prove; stage
timings and the verification equation in the exponent.
usage: python tools/qap_large.py [log_k] [reps]"""
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interactive_zkp_study_b200 import native as nat  # noqa: E402
from interactive_zkp_study_b200.zkp.groth16 import qap_device as qd  # noqa: E402

R = nat.R_MOD


def circuit(k, seed):
    """k gates over k + 2 wires (wire 0 = 1, wire 1 = input): products and linear combinations of earlier wires."""
    rng = random.Random(seed)
    w = [1, rng.randrange(2, 1 << 20)]
    ra, rb, rc = [], [], []
    for g in range(k):
        a, b = rng.randrange(len(w)), rng.randrange(len(w))
        if rng.random() < 0.5:
            ra.append({a: 1})
            rb.append({b: 1})
            w.append(w[a] * w[b] % R)
        else:
            cst = rng.randrange(1, 1 << 16)
            row = {a: 1}
            row[b] = (row.get(b, 0) + cst) % R
            ra.append(row)
            rb.append({0: 1})
            w.append((w[a] + cst * w[b]) % R)
        rc.append({g + 2: 1})
    return ra, rb, rc, w


def run(log_k=20, reps=3, quiet=False):
    k = 1 << log_k
    m = k + 2
    t0 = time.perf_counter()
    ra, rb, rc, wit = circuit(k, 99)
    r1cs = qd.SparseR1CS.from_rows(k, m, ra, rb, rc)
    gen_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    dev = qd.DeviceR1CS(r1cs)
    rng = random.Random(7)
    alpha, beta, gamma, delta, x = (rng.randrange(1, R) for _ in range(5))
    keys = qd.setup(dev, alpha, beta, gamma, delta, x, lcm=True, precompute=True)
    setup_s = time.perf_counter() - t0
    w = nat.scalars_load(nat.fr_vec_bytes(wit), m)
    r, s = rng.randrange(R), rng.randrange(R)
    # stage timing: witness polynomials alone, then the whole prove
    wp, pr = [], []
    for _ in range(reps + 1):
        nat.sync()
        nat.timer_start()
        us = qd.witness_polys(dev, w, True)
        wp.append(nat.timer_stop())
        for h in us:
            h.free()
        nat.sync()
        nat.timer_start()
        A, B, C = qd.prove(keys, dev, w, r, s)
        pr.append(nat.timer_stop())
    A, B, C, hq, hr, uA, uB, uC = qd.prove(keys, dev, w, r, s, keep=True)
    ev = lambda h, n: nat.fr_poly_eval_dev(h, 0, n, x)
    ax, bx, cx, zx, hx = ev(uA, k), ev(uB, k), ev(uC, k), ev(keys.Z, k + 1), ev(hq, k - 1)
    divisible = nat.scalars_is_zero(hr, 0, k) and (ax * bx - cx) % R == hx * zx % R
    from oracle import bn254
    a = (alpha + ax + r * delta) % R
    b = (beta + bx + s * delta) % R
    # public wires 0 and 1: val_i from the Lagrange basis at x (host, two sparse columns)
    sA, sB, sC = qd._scales(k, True)
    lag = nat.fr_vec_from_bytes(nat.scalars_download(nat.fr_ap_lagrange_dev(k, x), 0, k))
    pub_term = 0
    for i in keys.pub_idx:
        ai = sum(row[i] * lag[g] for g, row in enumerate(ra) if i in row) % R * sA
        bi = sum(row[i] * lag[g] for g, row in enumerate(rb) if i in row) % R * sB
        ci = sum(row[i] * lag[g] for g, row in enumerate(rc) if i in row) % R * sC
        pub_term += wit[i] * (beta * ai + alpha * bi + ci)
    c = (a * b - alpha * beta - pub_term) % R * pow(delta, -1, R) % R
    ok = ((int(A[0]), int(A[1])) == bn254.g1_mul(bn254.G1, a)
          and ((int(B[0].coeffs[0]), int(B[0].coeffs[1])), (int(B[1].coeffs[0]), int(B[1].coeffs[1]))) == bn254.g2_mul(bn254.G2, b)
          and (int(C[0]), int(C[1])) == bn254.g1_mul(bn254.G1, c))
    res = {"constraints": k, "wires": m, "nnz": [int(len(ci)) for _, ci, _ in r1cs.mats],
           "prove_ms": min(pr[1:]), "witness_polys_ms": min(wp[1:]), "first_call_ms": pr[0],
           "circuit_gen_s": gen_s, "setup_s": setup_s, "quotient_remainder_zero": bool(divisible),
           "verification_equation_in_exponent": bool(ok),
           "work": "3 sparse mat-vecs + 3 interpolations on {1..k} (subproduct tree, batched NTTs, cached M side), "
                   "quotient by Z = prod (x - j), 2 G1 MSMs + 1 G2 MSM"}
    if not quiet:
        print(res)
    return res


if __name__ == "__main__":
    run(int(sys.argv[1]) if len(sys.argv) > 1 else 20, int(sys.argv[2]) if len(sys.argv) > 2 else 3)
