#!/bin/bash
mkdir -p gpurun_out
for prio in 0 -1; do
  SPEI_BENCH_SEARCH_PRIORITY=$prio timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tmp.json 2> gpurun_out/bench_tmp.err; echo "prio $prio exit $?"
  python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_tmp.json'))
print({k:d[k] for k in ('value','ms_per_step')}, 'tc_ms', round(d['roofline']['kernel_ms'],3), 'e2e', round(d['e2e']['value'],1))
PY
done
