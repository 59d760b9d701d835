set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
python -m pytest tests/test_gpu_field.py tests/test_gpu_msm.py -x -q -m gpu 2>&1 | tail -30
python - <<'PY'
from interactive_zkp_study_b200 import native as nat
print(nat.device_info())
for v,name in [(0,'IMAD.WIDE'),(1,'IMAD.lo'),(2,'IMAD.HI'),(3,'Fp mul (136 MAC)')]:
    print(name, nat.imad_peak(v), 'G MAC/s')
from oracle import bn254
import time
for logn in (16,18,20):
    n=1<<logn
    s_h = nat.scalars_generate(0x5EED0002, n); k_h = nat.scalars_generate(0x5EED0001, n)
    t0=time.time(); table = nat.g1_fixed_base_mul_dev(nat.g1_bytes(bn254.G1), s_h, n); print('gen',logn,time.time()-t0)
    nat.g1_msm_dev(table,0,k_h,0,n)
    for rep in range(3):
        nat.timer_start(); nat.g1_msm_dev(table,0,k_h,0,n); ms=nat.timer_stop()
        print('msm 2^%d: %.3f ms  %.1f Mpts/s'%(logn,ms,n/ms/1e3))
PY
