#!/bin/bash
# A/B of library variants on ONE box: for each lib in build_ab/ (plus the default) run the search parity stage
# and two short benches; prints the tcgen05 kernel time of each.  Usage: tools/ab_libs.sh [variant ...]
mkdir -p gpurun_out
VARS="${@:-default k8 k8nw k16nw}"
for round in 1 2; do
for v in $VARS; do
  if [ "$v" = default ]; then unset SPEINET_B200_LIB; else export SPEINET_B200_LIB=$PWD/build_ab/lib_$v.so; fi
  if [ $round = 1 ]; then
    timeout 300 python tests/diag/gpu_diag.py --stage tc --out gpurun_out/ab_${v} > gpurun_out/ab_${v}_tc.log 2>&1; echo "$v tc exit $?"
    grep -E "^tc " gpurun_out/ab_${v}_tc.log | cut -c 1-260
  fi
  timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/ab_${v}_bench$round.json 2> gpurun_out/ab_${v}_bench$round.err; echo "$v bench exit $?"
  python - <<PY
import json
d=json.load(open('gpurun_out/ab_${v}_bench$round.json'))
print('$v', 'round $round', 'ms_per_step', round(d['ms_per_step'],3), 'tc_ms', round(d['roofline']['kernel_ms'],3), 'frac', round(d['roofline']['frac'],3), d['search_stats_last_step'], d['clocks'])
PY
done
done
