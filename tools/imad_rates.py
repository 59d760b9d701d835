"""Integer-multiply pipe rates of this GPU (zkp_imad_peak variants; see include/zkp_b200_diag.h).
usage: python tools/imad_rates.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interactive_zkp_study_b200 import native as nat  # noqa: E402

NAMES = {0: "IMAD.WIDE.U32[.X] carry chains (production form; the roofline peak)", 1: "IMAD (lo)", 2: "IMAD.HI.U32",
         3: "Fp Montgomery product chains (x136)", 4: "mad.wide.u32, register factors (IMAD.WIDE + IADD3 pair)",
         5: "variant 4 + 64-bit shift/add on the ALU, 1:1"}
info = nat.device_info()
print(info)
for v in range(6):
    g = nat.imad_peak(v)
    print("variant %d  %-72s %9.1f G limb-MAC/s  = %.1f lanes/clk/SM at 1.965 GHz" % (v, NAMES[v], g, g / info["sm_count"] / 1.965))
