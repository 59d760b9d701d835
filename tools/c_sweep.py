"""Window-width sweep for window-precomputed tables: best-of-3 MSM time per (n, c).
usage: python tools/c_sweep.py [n ...]   (n may be any integer, e.g. 3145726)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interactive_zkp_study_b200 import native as nat  # noqa: E402

G1 = (1).to_bytes(32, "little") + (2).to_bytes(32, "little")
sizes = [int(a) for a in sys.argv[1:]] or [1 << 14, 1 << 16, 1 << 18, 1 << 20]
for n in sizes:
    s_h = nat.scalars_generate(0x5EED0002, n)
    k_h = nat.scalars_generate(0x5EED0001, n)
    row = []
    lg = n.bit_length() - 1
    for c in (range(4, 15) if n <= 2048 else range(max(4, lg - 6), min(20, lg + 2) + 1)):
        table = nat.g1_fixed_base_mul_dev(G1, s_h, n)
        try:
            nat.table_precompute(table, c)
        except Exception as e:
            row.append((c, None))
            table.free()
            continue
        best = 1e9
        for _ in range(4):
            nat.timer_start()
            nat.g1_msm_dev(table, 0, k_h, 0, n)
            best = min(best, nat.timer_stop())
        row.append((c, round(best, 3)))
        table.free()
    good = [r for r in row if r[1] is not None]
    print("n=%d (2^%.2f): best c=%d %.3f ms | %s" % (n, __import__("math").log2(n), min(good, key=lambda r: r[1])[0],
                                                     min(r[1] for r in good), " ".join("%d:%s" % r for r in row)), flush=True)
    s_h.free()
    k_h.free()
