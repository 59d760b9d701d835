python -m pytest tests/test_gpu_msm.py -x -q -m gpu 2>&1 | tail -5
for lg in 16 18 20; do python tools/msm_once.py $lg 3; done
python tools/msm_once.py 20 2 > /dev/null && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1_v3.csv python tools/msm_once.py 20 2 > gpurun_out/ncu.log 2>&1
