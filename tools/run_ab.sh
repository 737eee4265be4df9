mkdir -p gpurun_out
for v in ${VARIANTS}; do
  echo "== $v"; SPEINET_B200_LIB=$PWD/build_ab/lib_$v.so timeout 40 python tools/time_candidates.py tcs 2>&1 | tail -n ${TAILN:-1}
  if [ -n "$CHECK" ]; then SPEINET_B200_LIB=$PWD/build_ab/lib_$v.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "search or certified or second_pass or module_matches" 2>&1 | tail -n 2; fi
done
echo "== default"; timeout 40 python tools/time_candidates.py tcs 2>&1 | tail -n 1
