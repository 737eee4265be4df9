#!/bin/bash
# Runs every diagnostics stage in its own process (a faulting stage must not poison the next).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/diag_smi.log 2>&1
for s in ${STAGES:-env fold exact fuse tile tc time720}; do
  timeout 400 python tests/diag/gpu_diag.py --stage $s > gpurun_out/diag_$s.log 2>&1
  echo "stage $s exit $?"
done
for s in ${STAGES:-env fold exact fuse tile tc time720}; do echo "=== $s"; tail -c 1500 gpurun_out/diag_$s.log; done
