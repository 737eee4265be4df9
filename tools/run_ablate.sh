#!/bin/bash
# candidates-only timing of library variants in build_ab/ (kernel experiments)
mkdir -p gpurun_out
: > gpurun_out/ablate.log
for v in default "$@"; do
  if [ "$v" = default ]; then unset SPEINET_B200_LIB; else export SPEINET_B200_LIB=$PWD/build_ab/lib_$v.so; fi
  echo "== $v" | tee -a gpurun_out/ablate.log
  timeout 120 python tools/time_candidates.py ${MODE:-tcs} 2>&1 | tail -n ${TAILN:-1} | tee -a gpurun_out/ablate.log
done
