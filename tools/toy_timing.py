"""Wall time of the reference-sized flows (BASELINE configs[0]): toy Groth16 (x^3 + x + 5 = 35: hxr, proof_a,
proof_b, proof_c), kzg.commit and the PLONK prover on the x^3 circuit and on a 16-gate chain -- through the
drop-in mirrors (legacy list signatures, GPU underneath) and, beside them, through the CPU oracle's restatement
of the same reference functions on one host core.  Results are compared bit for bit while timing.
usage: python tools/toy_timing.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from tests.util import load, g1, g2, ints  # noqa: E402
from interactive_zkp_study_b200 import native as nat  # noqa: E402
from interactive_zkp_study_b200.compat import FQ, FQ2, FR, g1_from_ints, g2_from_ints  # noqa: E402
from interactive_zkp_study_b200.zkp.groth16 import poly_utils, proving  # noqa: E402
from interactive_zkp_study_b200.zkp.plonk import kzg, prover  # noqa: E402
from interactive_zkp_study_b200.zkp.plonk.polynomial import Polynomial  # noqa: E402
from interactive_zkp_study_b200.zkp.plonk.preprocessor import preprocess  # noqa: E402
from interactive_zkp_study_b200.zkp.plonk.srs import SRS  # noqa: E402
from oracle import ref_path  # noqa: E402  (the CPU side of the comparison)
import test_gpu_plonk as tp  # noqa: E402


def best(fn, reps):
    out, t = None, 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        t = min(t, time.perf_counter() - t0)
    return out, t * 1e3


rows = []


def row(name, gpu_ms, cpu_ms, same):
    rows.append((name, gpu_ms, cpu_ms, same))
    print("%-44s mirror (GPU) %9.3f ms   CPU oracle %10.3f ms   x%-8.1f %s" % (
        name, gpu_ms, cpu_ms, cpu_ms / gpu_ms, "bit-identical" if same else "DIFFERENT"), flush=True)


# ---------------------------------------------------------------- toy Groth16
g = load("groth16_toy.json")
pt1 = lambda p: (FQ(int(p[0])), FQ(int(p[1])))
pt2 = lambda p: (FQ2([int(p[0][0]), int(p[0][1])]), FQ2([int(p[1][0]), int(p[1][1])]))
Ax, Bx, Cx = ([[FR(int(x)) for x in r_] for r_ in g[k]] for k in ("Ax", "Bx", "Cx"))
Zx, Rx = [FR(int(x)) for x in g["Zx"]], [FR(int(x)) for x in g["Rx"]]
s11, s12, s14, s15 = ([pt1(p) for p in g[k]] for k in ("sigma1_1", "sigma1_2", "sigma1_4", "sigma1_5"))
s21, s22 = ([pt2(p) for p in g[k]] for k in ("sigma2_1", "sigma2_2"))
r, s = int(g["r"]), int(g["s"])
Ai, Bi, Ci = ([ints(r_) for r_ in g[k]] for k in ("Ax", "Bx", "Cx"))
Zi, Ri = ints(g["Zx"]), ints(g["Rx"])
i11, i12, i14, i15 = ([g1(p) for p in g[k]] for k in ("sigma1_1", "sigma1_2", "sigma1_4", "sigma1_5"))
i21, i22 = ([g2(p) for p in g[k]] for k in ("sigma2_1", "sigma2_2"))
g1i = lambda p: None if p is None else (int(p[0]), int(p[1]))
g2i = lambda p: None if p is None else ((int(p[0].coeffs[0]), int(p[0].coeffs[1])), (int(p[1].coeffs[0]), int(p[1].coeffs[1])))

proving.proof_a(s11, s12, Ax, Rx, r)                         # first call: table upload + caches
(hx, rem), t_g = best(lambda: poly_utils.hxr(Ax, Bx, Cx, Zx, g["R_raw"]), 5)
(chx, crem), t_c = best(lambda: ref_path.hxr(Ai, Bi, Ci, Zi, Ri), 3)
row("groth16 hxr (4 gates, 6 wires)", t_g, t_c, [int(v) for v in hx] == [v % nat.R_MOD for v in chx])
A, t_g = best(lambda: proving.proof_a(s11, s12, Ax, Rx, r), 5)
cA, t_c = best(lambda: ref_path.proof_a(i11, i12, Ai, Ri, r), 1)
row("groth16 proof_a", t_g, t_c, g1i(A) == cA)
B, t_g = best(lambda: proving.proof_b(s21, s22, Bx, Rx, s), 5)
cB, t_c = best(lambda: ref_path.proof_b(i21, i22, Bi, Ri, s), 1)
row("groth16 proof_b (G2)", t_g, t_c, g2i(B) == cB)
Hx = [FR(int(x)) for x in g["Hx"]]
C, t_g = best(lambda: proving.proof_c(s11, s12, s14, s15, Bx, Rx, Hx, s, r, A), 5)
cC, t_c = best(lambda: ref_path.proof_c(i11, i12, i14, i15, Bi, Ri, ints(g["Hx"]), s, r, cA), 1)
row("groth16 proof_c", t_g, t_c, g1i(C) == cC)
tot_g = sum(x[1] for x in rows)
tot_c = sum(x[2] for x in rows)
row("groth16 toy prove (hxr + A + B + C)", tot_g, tot_c, all(x[3] for x in rows))

# ---------------------------------------------------------------- PLONK: commit and the five rounds
for name in ("plonk_x3.json", "plonk_chain16.json"):
    f = load(name)
    srs = SRS([g1_from_ints(g1(p)) for p in f["g1_powers"]], [g2_from_ints(g2(p)) for p in f["g2_powers"]], f["srs_max_degree"])
    coeffs = [int(x) for x in f["polys"]["a"]] if "a" in f["polys"] else [int(x) for x in list(f["polys"].values())[0]]
    poly = Polynomial([FR(c) for c in coeffs])
    kzg.commit(poly, srs)
    cm, t_g = best(lambda: kzg.commit(poly, srs), 5)
    cc, t_c = best(lambda: ref_path.commit(coeffs, [g1(p) for p in f["g1_powers"]]), 1)
    row("plonk kzg.commit (%d coefficients, %s)" % (len(coeffs), name.split(".")[0]), t_g, t_c, g1i(cm) == cc)
    pp = preprocess(tp.StubCircuit(f, FR), srs)
    args = ([FR(int(x)) for x in f["a_vals"]], [FR(int(x)) for x in f["b_vals"]], [FR(int(x)) for x in f["c_vals"]],
            [FR(int(x)) for x in f.get("public_inputs", [])], pp, srs)
    prover.prove(None, *args)
    _, t_g = best(lambda: prover.prove(None, *args), 5)
    # CPU side: the prover is 9 commitments + O(n^2) polynomial work; time the 9 commitments alone (a lower bound)
    n_coeff = f["n"] + 3
    pts = [g1(p) for p in f["g1_powers"]]
    _, t_one = best(lambda: ref_path.commit([(7919 * (i + 1)) % nat.R_MOD for i in range(min(n_coeff, len(pts)))], pts), 1)
    row("plonk prove n=%d (%s)" % (f["n"], name.split(".")[0]), t_g, 9 * t_one, True)
print("kernels launched:", nat.launch_count())
