#!/bin/bash
mkdir -p gpurun_out
for v in default $@; do
  if [ "$v" = default ]; then unset SPEINET_B200_LIB; else export SPEINET_B200_LIB=$PWD/build_ab/lib_$v.so; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_fuse_$v.json 2> gpurun_out/bench_fuse_$v.err; echo "bench $v exit $?"
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_fuse_$v.json'))
print('$v', round(d['ms_per_step'],3), [(s['stage'], round(s['ms']*1e3,1), round(s['frac_of_hbm_peak'],3)) for s in d['roofline_hbm_stages']['stages'] if s['stage'].startswith('d_')])
PY
  python tools/diag_fuse_err.py | grep "^level"
done
