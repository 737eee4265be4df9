#!/bin/bash
# bench in both schedules, no CPU baseline
mkdir -p gpurun_out
for mode in "" "--no-overlap"; do
  timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline $mode > gpurun_out/bench_tmp.json 2> gpurun_out/bench_tmp.err; echo "bench '$mode' exit $?"
  python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_tmp.json'))
print({k:d[k] for k in ('value','ms_per_step')}, 'tc_ms', round(d['roofline']['kernel_ms'],3), 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],1), d['clocks'], d['search_stats_last_step'], d['config'].get('schedule'))
PY
  tail -n 2 gpurun_out/bench_tmp.err
done
