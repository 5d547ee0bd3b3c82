python -m pytest tests/test_gpu_constraints.py tests/test_gpu_biomes.py -m gpu -q -x 2>&1 | tail -3
python tools/profile_biomes.py 65536 2>&1 | tail -12
