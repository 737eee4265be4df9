#!/bin/bash
# ncu evidence for the bench command (B200_PROFILING.md recipe): plain run first, then the launch
# list, then full captures of the dominant kernel and (KERNELS=...) of the secondary ones.
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-bf16 --no-other"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list exit $?"
if [ -z "$SKIP_FULL" ]; then
$CMD > gpurun_out/ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:relevance_tcs_kernel -s 3 -c 1 -o gpurun_out/prof_tcs $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture (tcgen05) exit $?"
fi
if [ -n "$KERNELS" ]; then
  $CMD > gpurun_out/ncu_plain3.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k "regex:$KERNELS" -s 6 -c 18 -o gpurun_out/prof_rest $CMD > gpurun_out/ncu_full_rest.log 2>&1
  echo "full capture (rest) exit $?"
fi
