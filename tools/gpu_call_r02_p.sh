set -x
bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/r02_ab_every.log
