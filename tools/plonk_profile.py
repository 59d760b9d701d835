"""Attribute the device PLONK prover's time to native entry points (each call followed by a sync)."""
import collections, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from interactive_zkp_study_b200 import native as nat
from interactive_zkp_study_b200.zkp.plonk import device_prover as dp
import plonk_synth
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = 1 << log_n
key, wit, tau = plonk_synth.device_setup(plonk_synth.chain_circuit(n, seed=7))
dp.prove(key, *wit)
acc = collections.defaultdict(lambda: [0, 0.0])
names = ["scalars_alloc", "scalars_copy", "ntt_dev", "vec_op_dev", "scalars_load", "g1_msm_dev", "plonk_perm_terms_dev",
         "batch_inverse_dev", "scan_dev", "scalars_convert", "plonk_quotient_dev", "scalars_is_zero", "fr_poly_eval_dev",
         "axpy_dev", "scalars_add_const", "div_linear_dev", "scalars_upload", "g1_msm_dev_batch", "fr_poly_eval_multi_dev",
         "lincomb_dev"]
orig = {k: getattr(nat, k) for k in names}
def wrap(k):
    f = orig[k]
    def g(*a, **kw):
        nat.sync(); t0 = time.perf_counter(); r = f(*a, **kw); nat.sync()
        acc[k][0] += 1; acc[k][1] += time.perf_counter() - t0
        return r
    return g
for k in names:
    setattr(nat, k, wrap(k))
t0 = time.perf_counter(); dp.prove(key, *wit); tot = time.perf_counter() - t0
for k, (c, t) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print("%-24s %4d calls %8.2f ms" % (k, c, t * 1e3))
print("sum %.2f ms, wall %.2f ms" % (sum(v[1] for v in acc.values()) * 1e3, tot * 1e3))
