"""Single-thread latency of the Fp inversion: batched division steps (Mont256::inv, csrc/modinv30.cuh) against the
binary extended Euclid it replaced (Mont256::inv_euclid)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interactive_zkp_study_b200 import native as nat  # noqa: E402

for mode, name in ((0, "Fp product"), (10, "inversion, division steps"), (11, "inversion, binary Euclid")):
    print("%-28s %10.1f ns" % (name, nat.latency_probe(mode)))
