"""Synthetic PLONK prove at 2^k gates (BASELINE config 4) on the device-resident prover; the proof is
checked by the oracle's restatement of the reference verifier (pairings on the CPU).
usage: python tools/plonk_large.py [log_n] [reps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from interactive_zkp_study_b200 import native as nat  # noqa: E402
from interactive_zkp_study_b200.zkp.plonk import device_prover as dp  # noqa: E402
import plonk_synth  # noqa: E402


def run(log_n=20, reps=3, verify=True, quiet=False, comm=None):
    """comm: a sharded.Communicator -> the SRS is sharded by point range and every commitment is a collective
    MSM over the ranks; every rank runs this function with the same arguments."""
    n = 1 << log_n
    t0 = time.perf_counter()
    circ = plonk_synth.chain_circuit(n, seed=7)
    gen_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    key, wit, tau = plonk_synth.device_setup(circ, comm=comm)
    setup_s = time.perf_counter() - t0
    times = []
    proof = None
    blinds = list(range(11, 20)) if comm is not None else None    # the ranks must draw the same blinding scalars
    for _ in range(reps + 1):
        nat.sync()
        if comm is not None:
            comm.barrier()
        nat.timer_start()
        proof = dp.prove(key, *wit, blinds=blinds)
        times.append(nat.timer_stop())
    ok = None
    if verify:
        from oracle import bn254, plonk_verifier
        pd = {k: ((int(v[0]), int(v[1])) if k.endswith("_comm") else int(v)) for k, v in vars(proof).items()}
        pre = {k: key.comm[k] for k in dp.CIRCUIT_POLYS}
        ok = plonk_verifier.verify(pd, pre, n, key.omega, [bn254.G2, bn254.g2_mul(bn254.G2, tau)])
    res = {"gates": n, "prove_ms": min(times[1:]), "first_call_ms": times[0], "circuit_gen_s": gen_s, "setup_s": setup_s,
           "accepted_by_oracle_verifier": ok, "quotient_coset": "%dn" % key.ext, "n_gpus": 1 if comm is None else comm.world,
           "work": "9 G1 MSMs of ~n points (window-precomputed SRS), 4 iNTT(n), 4 coset NTT + 1 coset iNTT of size 4n, "
                   "fused quotient kernel, batch inverse + product scan, 2 sum-scan openings, 7 Horner evaluations"}
    if not quiet:
        print(res)
    return res


if __name__ == "__main__":
    run(int(sys.argv[1]) if len(sys.argv) > 1 else 20, int(sys.argv[2]) if len(sys.argv) > 2 else 3)
