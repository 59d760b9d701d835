python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 | tee gpurun_out/bench_r1_b.json | cut -c1-900
