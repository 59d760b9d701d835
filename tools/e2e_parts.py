"""End-to-end G1 MSM (zkp_g1_msm_table: scalars in pinned host memory) against the number of point ranges of the
part-streamed form, beside the device-resident MSM.  usage: python tools/e2e_parts.py [log_n]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interactive_zkp_study_b200 import native as nat  # noqa: E402
from oracle import bn254  # noqa: E402

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = 1 << log_n
s_h = nat.scalars_generate(0x5EED0002, n)
k_h = nat.scalars_generate(0x5EED0001, n)
table = nat.g1_fixed_base_mul_dev(nat.g1_bytes(bn254.G1), s_h, n)
nat.table_precompute(table)
pinned = nat.PinnedBuffer(32 * n)
pinned.write(nat.scalars_download(k_h, 0, n))
want = nat.g1_msm_dev(table, 0, k_h, 0, n)
reps = 10
for parts in (1, 3, 4, 5, 6):
    nat.msm_set_option("parts", parts)
    for _ in range(3):
        got = nat.g1_msm_table(table, 0, pinned.addr, n)
    assert got == want, "part-streamed result differs"
    nat.timer_start()
    for _ in range(reps):
        nat.g1_msm_table(table, 0, pinned.addr, n)
    e2e = nat.timer_stop() / reps
    for _ in range(2):
        nat.g1_msm_dev(table, 0, k_h, 0, n)
    nat.timer_start()
    for _ in range(reps):
        nat.g1_msm_dev(table, 0, k_h, 0, n)
    res = nat.timer_stop() / reps
    print("2^%d parts=%d   e2e %.3f ms (%.1f Mpts/s)   resident %.3f ms (%.1f Mpts/s)" %
          (log_n, parts, e2e, n / e2e / 1e3, res, n / res / 1e3))
