set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 2>gpurun_out/bench_err.log | tee gpurun_out/bench_r1_a.json
tail -5 gpurun_out/bench_err.log
python bench.py --impl reference --steps 3 --warmup 1 | tee gpurun_out/bench_ref_r1_a.json
