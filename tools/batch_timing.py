"""Throughput of K back-to-back 2^k MSMs: one call each vs one pipelined batch call."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interactive_zkp_study_b200 import native as nat
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
K = 8
n = 1 << log_n
t = nat.g1_fixed_base_mul_dev(nat.g1_bytes((1, 2)), nat.scalars_generate(2, n), n)
nat.table_precompute(t)
ks = [nat.scalars_generate(100 + i, n) for i in range(K)]
single = [nat.g1_msm_dev(t, 0, k, 0, n) for k in ks]
assert nat.g1_msm_dev_batch(t, [(k, 0, 0, n) for k in ks]) == single
for _ in range(2):
    nat.timer_start(); [nat.g1_msm_dev(t, 0, k, 0, n) for k in ks]; a = nat.timer_stop()
    nat.timer_start(); nat.g1_msm_dev_batch(t, [(k, 0, 0, n) for k in ks]); b = nat.timer_stop()
    print("2^%d x %d: sequential %.3f ms/MSM (%.1f Mpts/s), pipelined batch %.3f ms/MSM (%.1f Mpts/s)"
          % (log_n, K, a / K, n * K / a / 1e3, b / K, n * K / b / 1e3))
