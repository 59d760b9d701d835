"""Stage timing of the 2^k Groth16 device prover (quotient / A / B / C / tiny)."""
import os, random, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interactive_zkp_study_b200 import native as nat
from interactive_zkp_study_b200.zkp.groth16 import device_prover as dp
R = nat.R_MOD
log_k = int(sys.argv[1]) if len(sys.argv) > 1 else 20
k = 1 << log_k; mp = k - 2
rng = random.Random(1)
alpha, beta, delta, x = (rng.randrange(1, R) for _ in range(4))
Z = nat.scalars_generate(0x5EED0300, k + 1); nat.scalars_upload(Z, k, nat.fe_bytes(1), 1)
zx = nat.fr_poly_eval_dev(Z, 0, k + 1, x)
priv = nat.fr_vec_from_bytes(nat.scalars_download(nat.scalars_generate(0x5EED0400, mp), 0, mp))
key = dp.setup_from_toxic(k, alpha, beta, delta, x, zx, priv)
uA, uB, uC = (nat.scalars_generate(0x5EED0100 + i, k) for i in range(3))
rx = nat.scalars_generate(0x5EED0200, mp)
def T(f, reps=3):
    f(); best = 1e9
    for _ in range(reps):
        nat.sync(); t0 = time.perf_counter(); f(); nat.sync(); best = min(best, time.perf_counter() - t0)
    return best * 1e3
def quot():
    hq, hr = nat.groth16_quotient_dev(uA, uB, uC, k, Z, k + 1); hq.free(); hr.free()
print("quotient %.2f ms" % T(quot))
sc = nat.scalars_alloc(k + 2)
print("msm A (G1, k+2) %.2f ms" % T(lambda: nat.g1_msm_dev(key.TA, 0, sc, 0, k + 2)))
nat.scalars_copy(sc, 0, uA, 0, k)
print("msm A random scalars %.2f ms" % T(lambda: nat.g1_msm_dev(key.TA, 0, sc, 0, k + 2)))
print("msm B (G2, k+2) %.2f ms" % T(lambda: nat.g2_msm_dev(key.TB2, 0, sc, 0, k + 2)))
nC = key.TC.n
scC = nat.scalars_generate(7, nC)
print("msm C (G1, %d) %.2f ms" % (nC, T(lambda: nat.g1_msm_dev(key.TC, 0, scC, 0, nC))))
one = nat.fr_vec_bytes([1]); G = nat.g1_bytes((1, 2))
print("tiny msm (2 pts) %.2f ms" % T(lambda: nat.g1_msm(G + G, nat.fe_bytes(12345) + one, 2)))
def allocs():
    a = nat.scalars_alloc(k + 2); b = nat.scalars_alloc(nC); a.free(); b.free()
print("2 allocs+frees %.2f ms" % T(allocs))
print("full prove %.2f ms" % T(lambda: dp.prove(key, uA, uB, uC, Z, rx, 5, 7)))
