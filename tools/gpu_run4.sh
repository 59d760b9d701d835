python -m pytest tests/test_gpu_ntt.py -x -q -m gpu 2>&1 | tail -15
python -m pytest tests/test_gpu_msm.py -x -q -m gpu 2>&1 | tail -3
ZKP_B200_TRACE=1 python tools/msm_once.py 20 2 2>&1 | tail -24
