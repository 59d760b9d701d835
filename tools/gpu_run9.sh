python -m pytest tests/test_gpu_field.py tests/test_gpu_msm.py -x -q -m gpu 2>&1 | tail -5
for c in 14 16 17 18 19 20; do echo "== precompute c=$c"; python tools/msm_once.py 20 3 $c | tail -2; done
ZKP_B200_TRACE=1 python tools/msm_once.py 20 2 17 2>&1 | tail -30
