"""A/B of the two bucket accumulations (XYZZ chains vs affine tree) on device-resident MSMs.
usage: python tools/tree_ab.py [g1|g2] [log_n] [precompute_window_bits, -1 = plain table]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interactive_zkp_study_b200 import native as nat  # noqa: E402
from oracle import bn254  # noqa: E402

group = sys.argv[1] if len(sys.argv) > 1 else "g1"
log_n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
pre = int(sys.argv[3]) if len(sys.argv) > 3 else 0
n = 1 << log_n
s_h = nat.scalars_generate(0x5EED0002, n)
k_h = nat.scalars_generate(0x5EED0001, n)
if group == "g1":
    table = nat.g1_fixed_base_mul_dev(nat.g1_bytes(bn254.G1), s_h, n)
    msm = nat.g1_msm_dev
else:
    table = nat.g2_fixed_base_mul_dev(nat.g2_bytes(bn254.G2), s_h, n)
    msm = nat.g2_msm_dev
if pre >= 0:
    nat.table_precompute(table, pre)
nat.msm_profile(True)
results = {}
variants = [("xyzz", 1, 0, 0)] + [("tree K=%d B=%d" % (k, b), 2, b, k) for k in (1, 2, 3, 4) for b in (16, 32, 64)]
for name, mode, items, rounds in variants:
    nat.msm_set_option("accumulate", mode)
    nat.msm_set_option("tree_items", items)
    nat.msm_set_option("tree_rounds", rounds)
    best, acc = 1e9, 0.0
    for _ in range(4):
        nat.timer_start()
        r = msm(table, 0, k_h, 0, n)
        ms = nat.timer_stop()
        if ms < best:
            best, acc = ms, nat.msm_last_profile("accumulate")
    results[name] = r
    print("%s 2^%d %-14s %8.3f ms  (accumulate stage %8.1f us)  %.1f Mpts/s" % (group, log_n, name, best, acc, n / best / 1e3))
assert len(set(map(str, results.values()))) == 1, "results differ between the accumulations"
print("all variants agree")
