# A/B timing of library variants in ONE call (box-to-box variance is +-2 %): hector_b200/ab_*.so
for rep in 1 2; do
for f in hector_b200/libhector_b200.so hector_b200/ab_*.so; do
  echo "== $f"; HECTOR_B200_LIB=$PWD/$f python tools/profile_run.py 65536 4 | grep "run ms" | tail -3 | tr '\n' ' '; echo
done
done
