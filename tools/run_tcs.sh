#!/bin/bash
# GPU check of the tap-sharing search: raw accumulator tile, search parity, robustness, bench A/B vs dense
mkdir -p gpurun_out
for s in ${STAGES:-tile_tcs tcs}; do timeout 300 python tests/diag/gpu_diag.py --stage $s > gpurun_out/diag_$s.log 2>&1; echo "stage $s exit $?"; grep -E "^(tcs|tile|robust) " gpurun_out/diag_$s.log | cut -c 1-600; done
for m in tcs tc; do
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --search $m > gpurun_out/bench_$m.json 2> gpurun_out/bench_$m.err; echo "bench $m exit $?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$m.json'))
print('$m', {k:d[k] for k in ('value','ms_per_step')}, d['roofline']['kernel_ms'], d['roofline']['frac'], d['clocks'], d['search_stats_last_step'])
PY
tail -n 3 gpurun_out/bench_$m.err
done
