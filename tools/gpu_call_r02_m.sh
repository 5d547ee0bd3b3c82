set -x
python -m pytest tests -m gpu -q > gpurun_out/r02_gputests_m.log 2>&1; tail -15 gpurun_out/r02_gputests_m.log
