#!/bin/bash
# full ncu capture of the tap-sharing tcgen05 kernel (one launch) on the bench command
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --search tcs"
$CMD > gpurun_out/ncu_tcs_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:relevance_tcs_kernel -s 3 -c 1 -o gpurun_out/prof_tcs $CMD > gpurun_out/ncu_tcs_full.log 2>&1
echo "full capture (tcs) exit $?"
