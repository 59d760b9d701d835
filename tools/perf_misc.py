"""Device-resident timings of the secondary kernels: Fr NTT (2^16..2^24) and G2 MSM."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interactive_zkp_study_b200 import native as nat  # noqa: E402
from interactive_zkp_study_b200 import _lib  # noqa: E402
import ctypes  # noqa: E402

R = nat.R_MOD


def omega(log_n):
    return pow(5, (R - 1) >> log_n, R)


def ntt_dev(h, log_n, inverse=False, shift=None):
    cs = nat.fe_bytes(shift) if shift is not None else None
    nat.check(_lib.lib().zkp_fr_ntt_dev(h.handle, 0, log_n, nat.buf(nat.fe_bytes(omega(log_n))), 1 if inverse else 0,
                                        nat.buf(cs)))


for log_n in [int(x) for x in (sys.argv[1:] or ["16", "20", "22", "24"])]:
    n = 1 << log_n
    h = nat.scalars_generate(0x5EED0004, n)
    ntt_dev(h, log_n)
    ntt_dev(h, log_n, True)
    best = 1e9
    for _ in range(5):
        nat.timer_start()
        ntt_dev(h, log_n)
        best = min(best, nat.timer_stop())
    macs = 68.0 * n * log_n
    print("ntt 2^%d: %.3f ms  %.2f Gelem/s  %.2f T MAC/s  %.0f GB/s(64 B/elem)" %
          (log_n, best, n / best / 1e6, macs / best / 1e9, 64.0 * n / best / 1e6))
    h.free()

G2 = nat.g2_bytes(((10857046999023057135944570762232829481370756359578518086990519993285655852781,
                    11559732032986387107991004021392285783925812861821192530917403151452391805634),
                   (8495653923123431417604973247489272438418190587263600148770280649306958101930,
                    4082367875863433681332203403145435568316851327593401208105741076214120093531)))
for log_n in (14, 18):
    n = 1 << log_n
    sc = nat.scalars_download(nat.scalars_generate(0x5EED0003, n), 0, n)
    t = nat.g2_fixed_base_mul(G2, sc, n)
    k = nat.scalars_generate(0x5EED0001, n)
    nat.g2_msm_dev(t, 0, k, 0, n)
    nat.timer_start()
    nat.g2_msm_dev(t, 0, k, 0, n)
    ms = nat.timer_stop()
    print("g2 msm 2^%d: %.3f ms  %.2f Mpts/s" % (log_n, ms, n / ms / 1e3))
