#!/bin/bash
# quick GPU check: correctness stages of the search + bench without the CPU baseline
mkdir -p gpurun_out
STAGES="${STAGES:-tile tc}"
for s in $STAGES; do timeout 300 python tests/diag/gpu_diag.py --stage $s > gpurun_out/diag_$s.log 2>&1; echo "stage $s exit $?"; grep -E "^(tc|exact|tile) " gpurun_out/diag_$s.log | cut -c 1-400; done
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['value'], d['clocks'], d['search_stats_last_step'])
PY
tail -n 3 gpurun_out/bench_quick.err
