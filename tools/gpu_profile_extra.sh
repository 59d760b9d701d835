# Extra ncu captures (round 1): the widest reduction level of the MSM tail and the G2 bucket accumulation.
# Each ncu run is preceded by a plain run of the same command that exited 0.
set -x
python tools/msm_once.py 20 2 20 > gpurun_out/plain_once2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:msm_ws_level -s 1 -c 1 -o gpurun_out/prof_ws_level \
    python tools/msm_once.py 20 2 20 > gpurun_out/ncu_ws.log 2>&1
python tools/g2_once.py > gpurun_out/plain_g2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:msm_accumulate -s 1 -c 1 -o gpurun_out/prof_g2_accumulate \
    python tools/g2_once.py > gpurun_out/ncu_g2.log 2>&1
ls -la gpurun_out/*.ncu-rep
