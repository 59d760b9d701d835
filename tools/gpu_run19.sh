ZKP_B200_BATCHED_AFFINE=1 python -m pytest tests/test_gpu_msm.py -x -q -m gpu 2>&1 | tail -3
echo "== baseline c=17"; python tools/msm_once.py 20 3 17 | tail -2
echo "== batched affine c=17"; ZKP_B200_BATCHED_AFFINE=1 python tools/msm_once.py 20 3 17 | tail -2
ZKP_B200_BATCHED_AFFINE=1 ZKP_B200_TRACE=1 python tools/msm_once.py 20 2 17 2>&1 | grep -E "accumulate|TOTAL" | tail -2
echo "== batched affine c=16"; ZKP_B200_BATCHED_AFFINE=1 python tools/msm_once.py 20 3 16 | tail -1
echo "== batched affine plain"; ZKP_B200_BATCHED_AFFINE=1 python tools/msm_once.py 20 3 0 | tail -1
