# ncu evidence for the round: launch list of the bench command + one full capture of the top kernel
set -x
python bench.py --steps 2 --warmup 3 --no-verify > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-verify > gpurun_out/ncu_bench.log 2>&1
python tools/msm_once.py 20 2 > gpurun_out/plain_once.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:msm_accumulate -s 1 -c 1 -o gpurun_out/prof_accumulate \
    python tools/msm_once.py 20 2 > gpurun_out/ncu_once.log 2>&1
ls -la gpurun_out/
