# ncu evidence for the round: launch list of the bench command + one full capture of the top kernel.
# Every ncu run is preceded by a plain run of the same command that exited 0.
set -x
python bench.py --steps 2 --warmup 3 --no-verify --no-extras > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-verify --no-extras > gpurun_out/ncu_bench.log 2>&1
python tools/msm_once.py 20 2 20 > gpurun_out/plain_once.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:msm_accumulate -s 1 -c 1 -o gpurun_out/prof_accumulate \
    python tools/msm_once.py 20 2 20 > gpurun_out/ncu_once.log 2>&1
python tools/ntt_once.py 20 > gpurun_out/ntt_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ntt_pass -s 2 -c 2 -o gpurun_out/prof_ntt \
    python tools/ntt_once.py 20 > gpurun_out/ncu_ntt.log 2>&1
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_final.json 2>/dev/null
ls -la gpurun_out/ | tail -8
