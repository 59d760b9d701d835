"""Experiment: batched-affine pairwise additions vs the XYZZ mixed addition (adds per second)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interactive_zkp_study_b200 import native as nat, _lib  # noqa: E402
from oracle import bn254  # noqa: E402

G1 = (1).to_bytes(32, "little") + (2).to_bytes(32, "little")
n = 1 << int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
s_h = nat.scalars_generate(0x5EED0002, n)
table = nat.g1_fixed_base_mul_dev(G1, s_h, n)
pts = [nat.g1_from_bytes(nat.table_download(table, i, 1)) for i in range(8)]
for batch in (8, 16, 32):
    ms = ctypes.c_double()
    out = bytearray(64 * 4)
    nat.check(_lib.lib().zkp_dbg_affine_pairs(table.handle, batch, ctypes.byref(ms), nat.buf(out), 4))
    ok = all(nat.g1_from_bytes(bytes(out[64 * i:64 * i + 64])) == bn254.g1_add(pts[2 * i], pts[2 * i + 1]) for i in range(4))
    print("batch %2d: %.3f ms for %d additions = %.2f G add/s  correct=%s" % (batch, ms.value, n // 2, n / 2 / ms.value / 1e6, ok))
