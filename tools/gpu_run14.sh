python -m pytest tests/test_gpu_msm.py tests/test_gpu_groth16_large.py -x -q -m gpu 2>&1 | tail -3
python - <<'PY'
import time, sys, os
sys.path.insert(0, os.getcwd())
from interactive_zkp_study_b200 import native as nat
G1 = nat.g1_bytes((1, 2))
for lg in (20, 24):
    n = 1 << lg
    s = nat.scalars_generate(5, n)
    nat.sync(); t0 = time.perf_counter()
    t = nat.g1_fixed_base_mul_dev(G1, s, n)
    print("fixed-base 2^%d: %.1f ms  (%.1f Mpts/s)" % (lg, (time.perf_counter() - t0) * 1e3, n / (time.perf_counter() - t0) / 1e6))
    t.free(); s.free()
PY
