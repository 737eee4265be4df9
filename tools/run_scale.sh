#!/bin/bash
# 1/2/4/8-GPU weak-scaling bench on ONE box (launched exactly as the driver does), plus the reference arm at N=1
mkdir -p gpurun_out
for n in ${NS:-1 2 4 8}; do
  if [ $n = 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 $EXTRA > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $n --steps 20 --warmup 3 $EXTRA > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
  fi
  echo "n=$n exit $?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_n$n.json') if l.startswith('{')][-1])
    print($n, round(d['value'],1), 'fps', round(d['ms_per_step'],3), 'ms  e2e', round(d['e2e']['value'],1), d['clocks'])
except Exception as e: print('parse failed', e)
PY
done
if [ -z "$SKIP_REF" ]; then
  timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err
  echo "reference arm exit $?"; tail -c 600 gpurun_out/bench_ref_n1.json
fi
