"""Runs a few device-resident G1 MSMs of 2^k synthetic points (profiling driver for ncu).
usage: python tools/msm_once.py <log_n> [reps] [precompute_window_bits]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interactive_zkp_study_b200 import native as nat  # noqa: E402

G1 = (1).to_bytes(32, "little") + (2).to_bytes(32, "little")
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
pre = int(sys.argv[3]) if len(sys.argv) > 3 else 0
n = 1 << log_n
s_h = nat.scalars_generate(0x5EED0002, n)
k_h = nat.scalars_generate(0x5EED0001, n)
table = nat.g1_fixed_base_mul_dev(G1, s_h, n)
if pre:
    t0 = time.time()
    nat.table_precompute(table, pre)
    print("precompute c=%d: %.1f ms" % (pre, (time.time() - t0) * 1e3))
for _ in range(reps):
    nat.timer_start()
    r = nat.g1_msm_dev(table, 0, k_h, 0, n)
    ms = nat.timer_stop()
    print("msm 2^%d: %.3f ms  %.1f Mpts/s" % (log_n, ms, n / ms / 1e3), r[0] % 1000)
