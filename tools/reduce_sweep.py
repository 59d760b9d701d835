"""Sweep of the weighted-sum recursion's tunables (radix of the wide levels, the item counts at which the wide and
the quad-fused forms take over) on a device-resident G1 MSM of 2^k points, precomputed table and plain table.
usage: python tools/reduce_sweep.py [log_n]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interactive_zkp_study_b200 import native as nat  # noqa: E402

G1 = (1).to_bytes(32, "little") + (2).to_bytes(32, "little")
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = 1 << log_n
s_h = nat.scalars_generate(0x5EED0002, n)
k_h = nat.scalars_generate(0x5EED0001, n)


def best_of(table, reps=6):
    best, r = 1e9, None
    for _ in range(reps):
        nat.timer_start()
        r = nat.g1_msm_dev(table, 0, k_h, 0, n)
        best = min(best, nat.timer_stop())
    return best, r


for pre in (1, 0):
    table = nat.g1_fixed_base_mul_dev(G1, s_h, n)
    if pre:
        nat.table_precompute(table, 0)
    base, ref = best_of(table)
    print("%s table 2^%d, defaults: %.3f ms" % ("precomputed" if pre else "plain", log_n, base))
    for radix in (4, 8, 16):
        for wide in (15, 17, 19):
            for quad in (12, 14, 16):
                nat.msm_set_option("reduce_radix", radix)
                nat.msm_set_option("wide_log2", wide)
                nat.msm_set_option("quad_log2", quad)
                ms, r = best_of(table, 4)
                assert r == ref, (radix, wide, quad)
                print("  radix %2d  wide 2^%d  quad 2^%d : %.3f ms (%+.3f)" % (radix, wide, quad, ms, ms - base))
    for name in ("reduce_radix", "wide_log2", "quad_log2"):
        nat.msm_set_option(name, 0)
    table.free()
