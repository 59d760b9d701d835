"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck)."""
import os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
from interactive_zkp_study_b200 import native as nat
from oracle import bn254, ref_path
g.smoke()
rng = random.Random(3)
R = bn254.R
n = 600
s = nat.scalars_generate(11, n)
t = nat.g1_fixed_base_mul_dev(nat.g1_bytes(bn254.G1), s, n)
k = nat.scalars_generate(12, n)
plain = nat.g1_msm_dev(t, 0, k, 0, n)
nat.table_precompute(t, 7)
assert nat.g1_msm_dev(t, 0, k, 0, n) == plain
assert nat.g1_msm_dev(t, 100, k, 3, 333) is not None
a = [rng.randrange(R) for _ in range(300)]; b = [rng.randrange(R) for _ in range(77)]
q, r = nat.fr_poly_divmod(nat.fr_vec_bytes(a), 300, nat.fr_vec_bytes(b), 77)
wq, wr = ref_path.poly_div(a, b)
assert nat.fr_vec_from_bytes(q) == wq + [0] * (224 - len(wq))
assert nat.fr_vec_from_bytes(nat.fr_batch_inverse(nat.fr_vec_bytes(a), 300)) == [ref_path.inv(x) for x in a]
pp = nat.fr_vec_from_bytes(nat.fr_prefix_product(nat.fr_vec_bytes(a[:50]), 50))
acc, want = 1, []
for x in a[:50]:
    want.append(acc); acc = acc * x % R
assert pp == want
w = ref_path.get_root_of_unity(2048)
v = [rng.randrange(R) for _ in range(2048)]
assert nat.fr_vec_from_bytes(nat.fr_ntt(nat.fr_vec_bytes(v), 11, w, False, 5)) == ref_path.coset_fft(v, w, 5)
print("sanitize_small ok")
