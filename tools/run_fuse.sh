#!/bin/bash
# GPU check of the fusion kernel: parity tests, then stage timings with the tcgen05 and the mma.sync version
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "fus or decode or model_forward or install" 2>&1 | tail -5
for v in tc sync; do
  if [ $v = sync ]; then export SPEI_FUSE_MMA_SYNC=1; else unset SPEI_FUSE_MMA_SYNC; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_fuse_$v.json 2> gpurun_out/bench_fuse_$v.err; echo "bench $v exit $?"
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_fuse_$v.json'))
print('$v', round(d['ms_per_step'],3), [(s['stage'], round(s['ms']*1e3,1), round(s['frac_of_hbm_peak'],3)) for s in d['roofline_hbm_stages']['stages'] if s['stage'].startswith('d_')])
PY
  tail -n 3 gpurun_out/bench_fuse_$v.err
done
