#!/bin/bash
# GPU check of the fusion kernel: parity tests, then the stage timings of the bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "fus or decode or model_forward or install or pipeline" 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_fuse.json 2> gpurun_out/bench_fuse.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_fuse.json'))
print(round(d['ms_per_step'],3), [(s['stage'], round(s['ms']*1e3,1), round(s['frac_of_hbm_peak'],3)) for s in d['roofline_hbm_stages']['stages']])
PY
tail -n 3 gpurun_out/bench_fuse.err
