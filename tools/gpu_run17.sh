python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r1_c.json 2> gpurun_out/bench_r1_c.err; tail -3 gpurun_out/bench_r1_c.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r1_c.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['whole_msm']['frac'], d['clocks'])
print(json.dumps(d['extras'])[:1800])
PY
