mkdir -p gpurun_out
for v in ${VARIANTS}; do
  echo "== $v"; SPEINET_B200_LIB=$PWD/build_ab/lib_$v.so timeout 120 python tools/time_candidates.py tcs 2>&1 | tail -n 12
done
echo "== default"; timeout 120 python tools/time_candidates.py tcs 2>&1 | tail -n 3
