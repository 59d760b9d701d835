"""All single-thread / single-quad latency probes of the diagnostics ABI (zkp_latency_probe), ns per operation."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interactive_zkp_study_b200 import native as nat  # noqa: E402

NAMES = ["1 Fp product per step", "2 independent Fp products per step", "4 independent Fp products per step",
         "XYZZ add, inlined products", "XYZZ add, out-of-line products", "XYZZ mixed add", "XYZZ double",
         "quad XYZZ add, inlined products", "quad XYZZ add, out-of-line products", "quad XYZZ double",
         "Fp inversion, division steps", "Fp inversion, binary Euclid",
         "quad add without its edge-case tail", "quad add: the 4 select/product/broadcast levels only",
         "4 dependent Fp products, lone thread", "4 dependent Fp products, each broadcast"]
for mode, name in enumerate(NAMES):
    print("mode %2d  %-40s %10.1f ns" % (mode, name, nat.latency_probe(mode)))
