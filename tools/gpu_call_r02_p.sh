set -x
python -m pytest tests/test_gpu_biomes.py -q -x 2>&1 | tail -3
for f in hector_b200/libhector_b200.so hector_b200/ab_head.so; do
echo "== $f"; HECTOR_B200_LIB=$PWD/$f python tools/profile_biomes.py 65536
done 2>&1 | tee gpurun_out/r02_ab_biome_smem.log
