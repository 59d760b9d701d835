python -m pytest tests/test_gpu_ntt.py tests/test_gpu_poly.py tests/test_gpu_plonk_device.py tests/test_gpu_groth16_large.py -x -q -m gpu 2>&1 | tail -3
python tools/perf_misc.py 20 22 2>&1 | head -2
python tools/plonk_profile.py 20 2>&1 | head -4
