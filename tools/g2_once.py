"""Device-resident G2 MSM of 2^k points on a window-precomputed table (profiling / tuning driver)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interactive_zkp_study_b200 import native as nat
G2 = nat.g2_bytes(((10857046999023057135944570762232829481370756359578518086990519993285655852781,
                    11559732032986387107991004021392285783925812861821192530917403151452391805634),
                   (8495653923123431417604973247489272438418190587263600148770280649306958101930,
                    4082367875863433681332203403145435568316851327593401208105741076214120093531)))
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = 1 << log_n
t = nat.g2_fixed_base_mul_dev(G2, nat.scalars_generate(3, n), n)
nat.table_precompute(t)
k = nat.scalars_generate(1, n)
for _ in range(3):
    nat.timer_start(); r = nat.g2_msm_dev(t, 0, k, 0, n); ms = nat.timer_stop()
    print("g2 msm 2^%d: %.3f ms  %.1f Mpts/s" % (log_n, ms, n / ms / 1e3))
