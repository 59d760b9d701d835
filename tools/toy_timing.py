"""Wall time of the reference-sized flows through the drop-in mirror (toy Groth16, PLONK x^3+x+5)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.util import load, g1, g2
from interactive_zkp_study_b200 import native as nat
from interactive_zkp_study_b200.compat import FQ, FQ2, FR, g1_from_ints, g2_from_ints
from interactive_zkp_study_b200.zkp.plonk.srs import SRS
from interactive_zkp_study_b200.zkp.plonk.preprocessor import preprocess
from interactive_zkp_study_b200.zkp.plonk import prover
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_plonk as tp
f = load("plonk_chain16.json")
srs = SRS([g1_from_ints(g1(p)) for p in f["g1_powers"]], [g2_from_ints(g2(p)) for p in f["g2_powers"]], f["srs_max_degree"])
pp = preprocess(tp.StubCircuit(f, FR), srs)
args = ([FR(int(x)) for x in f["a_vals"]], [FR(int(x)) for x in f["b_vals"]], [FR(int(x)) for x in f["c_vals"]], [], pp, srs)
prover.prove(None, *args)
t0 = time.perf_counter()
for _ in range(5):
    prover.prove(None, *args)
print("PLONK prove n=16 through the list-level mirror: %.1f ms" % ((time.perf_counter() - t0) / 5 * 1e3))
