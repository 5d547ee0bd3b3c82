# A/B of run-time switches in ONE call
for rep in 1 2; do
for v in 1 0; do
  echo "== HX_L2_PERSIST=$v"; HX_L2_PERSIST=$v python tools/profile_run.py 65536 4 | grep "run ms" | tail -3 | tr '\n' ' '; echo
done
done
