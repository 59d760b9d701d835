python -m pytest tests/test_gpu_msm.py -x -q -m gpu 2>&1 | tail -5
python tools/msm_once.py 20 3 0
ZKP_B200_TRACE=1 python tools/msm_once.py 20 3 16 2>&1 | tail -32
python tools/msm_once.py 20 3 15
python tools/msm_once.py 20 3 14
