set -x
python -m pytest tests/test_gpu_tracking.py tests/test_gpu_parity_at_size.py tests/test_gpu_constraints.py -q -x 2>&1 | tail -4
python tools/profile_tracked.py 65536 | tr '\n' ' '; echo
