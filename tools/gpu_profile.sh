#!/bin/bash
# ncu evidence of the round, one capture per gpurun call (the pool allows one ncu run per call):
#   tools/gpu_profile.sh launches | accumulate | g2 | ntt | ws
# Every ncu run is preceded by a plain run of the same command that exited 0.
set -x
BENCH="python bench.py --steps 2 --warmup 3 --no-extras --strong-log-n 0 --no-cpu"
case "$1" in
  launches)
    $BENCH > gpurun_out/plain_bench.log 2>&1 && \
    ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_bench.csv \
        $BENCH > gpurun_out/ncu_bench.log 2>&1 ;;
  accumulate)
    python tools/msm_once.py 20 2 20 > gpurun_out/plain_once.log 2>&1 && \
    ncu --set full --clock-control none --import-source on -k regex:msm_accumulate -s 1 -c 1 -o gpurun_out/prof_accumulate \
        python tools/msm_once.py 20 2 20 > gpurun_out/ncu_once.log 2>&1 ;;
  ws)
    python tools/msm_once.py 20 2 20 > gpurun_out/plain_once2.log 2>&1 && \
    ncu --set full --clock-control none --import-source on -k regex:msm_ws_level -s 1 -c 1 -o gpurun_out/prof_ws_level \
        python tools/msm_once.py 20 2 20 > gpurun_out/ncu_ws.log 2>&1 ;;
  g2)
    python tools/g2_once.py > gpurun_out/plain_g2.log 2>&1 && \
    ncu --set full --clock-control none --import-source on -k regex:msm_accumulate -s 1 -c 1 -o gpurun_out/prof_g2_accumulate \
        python tools/g2_once.py > gpurun_out/ncu_g2.log 2>&1 ;;
  ntt)
    python tools/ntt_once.py 20 > gpurun_out/ntt_plain.log 2>&1 && \
    ncu --set full --clock-control none --import-source on -k regex:ntt_pass -s 2 -c 2 -o gpurun_out/prof_ntt \
        python tools/ntt_once.py 20 > gpurun_out/ncu_ntt.log 2>&1 ;;
  *) echo "usage: $0 launches|accumulate|g2|ntt|ws"; exit 2 ;;
esac
ls -la gpurun_out/ | tail -5
