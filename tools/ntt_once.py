"""A few device-resident forward NTTs of 2^k elements (profiling driver for ncu); plain, coset and inverse coset."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interactive_zkp_study_b200 import native as nat
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = 1 << log_n
w = pow(5, (nat.R_MOD - 1) >> log_n, nat.R_MOD)
h = nat.scalars_generate(0x5EED0004, n)
for name, kw in (("plain", {}), ("coset", {"coset_shift": 5}), ("inverse coset", {"inverse": True, "coset_shift": 5})):
    best = 1e9
    for _ in range(4):
        nat.timer_start(); nat.ntt_dev(h, 0, log_n, w, **kw); best = min(best, nat.timer_stop())
    print("ntt 2^%d %-14s %.3f ms" % (log_n, name, best))
