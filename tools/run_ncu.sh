#!/bin/bash
# ncu evidence for the bench command (B200_PROFILING.md recipe): plain run first, then the launch
# list, then one full capture of the dominant kernel.
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 60 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:relevance_tc_kernel -s 3 -c 1 -o gpurun_out/prof_tc $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
tail -n 3 gpurun_out/ncu_launch.log gpurun_out/ncu_full.log
