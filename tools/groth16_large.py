"""Synthetic Groth16 prove at 2^k constraints (BASELINE config 3): device-resident key with known
toxic waste, timed stage by stage, every proof element checked against its closed-form discrete log.
usage: python tools/groth16_large.py [log_k] [reps]"""
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interactive_zkp_study_b200 import native as nat  # noqa: E402
from interactive_zkp_study_b200.zkp.groth16 import device_prover as dp  # noqa: E402


def run(log_k=20, reps=3, verify=True, quiet=False, comm=None):
    """comm: a sharded.Communicator -> the three MSMs of every proof are sharded by point range over its ranks
    (device_prover.prove_sharded); every rank runs this function with the same arguments."""
    R = nat.R_MOD
    k = 1 << log_k
    mp = k - 2
    rng = random.Random(2026)
    alpha, beta, delta, x = (rng.randrange(1, R) for _ in range(4))
    t0 = time.perf_counter()
    Z = nat.scalars_generate(0x5EED0300, k + 1)
    nat.scalars_upload(Z, k, nat.fe_bytes(1), 1)                      # monic
    zx = nat.fr_poly_eval_dev(Z, 0, k + 1, x)
    priv_h = nat.scalars_generate(0x5EED0400, mp)
    priv = nat.fr_vec_from_bytes(nat.scalars_download(priv_h, 0, mp))
    if comm is None:
        key = dp.setup_from_toxic(k, alpha, beta, delta, x, zx, priv, precompute=True)
        prove = lambda: dp.prove(key, uA, uB, uC, Z, rx, r, s)
    else:
        key = dp.setup_from_toxic_sharded(comm, k, alpha, beta, delta, x, zx, priv, precompute=True)
        prove = lambda: dp.prove_sharded(comm, key, uA, uB, uC, Z, rx, r, s)
    setup_s = time.perf_counter() - t0
    uA, uB, uC = (nat.scalars_generate(0x5EED0100 + i, k) for i in range(3))
    rx = nat.scalars_generate(0x5EED0200, mp)
    r, s = rng.randrange(R), rng.randrange(R)
    times = []
    for _ in range(reps + 1):
        nat.sync()
        if comm is not None:
            comm.barrier()
        nat.timer_start()
        A, B, C = prove()
        times.append(nat.timer_stop())
    ok = None
    if verify:
        from oracle import bn254
        ev = lambda h, n, at: nat.fr_poly_eval_dev(h, 0, n, at)
        if comm is None:
            A, B, C, hq, hr = dp.prove(key, uA, uB, uC, Z, rx, r, s, keep_quotient=True)
        else:
            hq, hr = nat.groth16_quotient_dev(uA, uB, uC, k, Z, k + 1, want_remainder=True)
        a = (alpha + ev(uA, k, x) + r * delta) % R
        b = (beta + ev(uB, k, x) + s * delta) % R
        rx_host = nat.fr_vec_from_bytes(nat.scalars_download(rx, 0, mp))
        wires = sum(p * q for p, q in zip(rx_host, priv)) % R
        c = (s * a + r * (beta + ev(uB, k, x)) + wires + ev(hq, k - 1, x) * zx % R * pow(delta, -1, R)) % R
        t = 0xfeedface12345
        ident = (ev(uA, k, t) * ev(uB, k, t) - ev(uC, k, t)) % R == (ev(hq, k - 1, t) * ev(Z, k + 1, t) + ev(hr, k, t)) % R
        ok = ((int(A[0]), int(A[1])) == bn254.g1_mul(bn254.G1, a)
              and ((int(B[0].coeffs[0]), int(B[0].coeffs[1])), (int(B[1].coeffs[0]), int(B[1].coeffs[1]))) == bn254.g2_mul(bn254.G2, b)
              and (int(C[0]), int(C[1])) == bn254.g1_mul(bn254.G1, c) and ident)
    best = min(times[1:])
    res = {"constraints": k, "private_wires": mp, "prove_ms": best, "first_call_ms": times[0], "setup_s": setup_s,
           "verified_against_discrete_logs": ok,
           "n_gpus": 1 if comm is None else comm.world,
           "work": "quotient h = (uA*uB - uC) div Z (NTT products + cached Newton inverse), 2 G1 MSMs (k+2, 3k points), "
                   "1 G2 MSM (k+2 points), no scalar multiplication (C is an independent MSM)" +
                   ("" if comm is None else "; each MSM sharded by point range over the ranks (zkp_g1/g2_msm_multi), the quotient "
                                            "computed by every rank (single-GPU NTT work)")}
    if not quiet:
        print(res)
    return res


if __name__ == "__main__":
    run(int(sys.argv[1]) if len(sys.argv) > 1 else 20, int(sys.argv[2]) if len(sys.argv) > 2 else 3)
