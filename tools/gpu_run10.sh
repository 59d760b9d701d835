python -m pytest tests/test_gpu_msm.py -x -q -m gpu 2>&1 | tail -3
for c in 15 16 17 18; do echo "== precompute c=$c"; python tools/msm_once.py 20 3 $c | tail -2; done
ZKP_B200_TRACE=1 python tools/msm_once.py 20 2 16 2>&1 | tail -26
