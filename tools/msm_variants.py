"""Engine variants of one device-resident G1 MSM of 2^k synthetic points on the window-precomputed table:
sort = counting sort with global atomics (1) / radix partition (2), split = 1..4 bucket-range parts.
Prints the best time of each and, for the automatic setting, the per-stage split (ZKP_B200_TRACE view).
usage: python tools/msm_variants.py [log_n] [reps] [precompute_window_bits]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interactive_zkp_study_b200 import native as nat  # noqa: E402
from oracle import bn254  # noqa: E402  (checker only)

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
pre = int(sys.argv[3]) if len(sys.argv) > 3 else 0
n = 1 << log_n
G1 = nat.g1_bytes((1, 2))
s_h = nat.scalars_generate(0x5EED0002, n)
ks = [nat.scalars_generate(0x5EED0001 + 0x1000 * v, n) for v in range(3)]
table = nat.g1_fixed_base_mul_dev(G1, s_h, n)
c = nat.table_precompute(table, pre)
want = [bn254.g1_mul(bn254.G1, nat.fr_dot_dev(k, 0, s_h, 0, n)) for k in ks]
print("2^%d points, window bits %d" % (log_n, c))
STAGES = ("digits+scan", "scatter", "tasks", "accumulate", "fold", "ws wide", "ws2", "ws fused", "horner")
for sort in (1, 2):
    for split in (1, 2, 3, 4):
        nat.msm_set_option("sort", sort)
        nat.msm_set_option("split", split)
        best = 1e9
        for i in range(reps + 1):
            nat.timer_start()
            got = nat.g1_msm_dev(table, 0, ks[i % 3], 0, n)
            ms = nat.timer_stop()
            assert got == want[i % 3], (sort, split)
            if i:
                best = min(best, ms)
        print("sort %d split %d: %.3f ms  %.1f Mpts/s" % (sort, split, best, n / best / 1e3), flush=True)
nat.msm_set_option("sort", 0)
nat.msm_set_option("split", 0)
os.environ["ZKP_B200_TRACE"] = "0"
nat.msm_profile(True)
for split in (1, 0):
    nat.msm_set_option("split", split)
    nat.g1_msm_dev(table, 0, ks[0], 0, n)
    nat.g1_msm_dev(table, 0, ks[1], 0, n)
    print("split %d stages (us):" % split, {s: round(nat.msm_last_profile(s), 1) for s in ("tasks", "accumulate", "horner")},
          "total", round(nat.msm_last_profile(None), 1))
