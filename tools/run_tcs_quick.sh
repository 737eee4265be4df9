#!/bin/bash
# tcs kernel iteration loop: parity (diag stages + pytest), candidates-only timing, robustness on smooth features
mkdir -p gpurun_out
for s in ${STAGES:-tile_tcs tcs robust_tcs}; do timeout 300 python tests/diag/gpu_diag.py --stage $s > gpurun_out/diag_$s.log 2>&1; echo "stage $s exit $?"; grep -E "^(tcs|tile|robust) " gpurun_out/diag_$s.log | cut -c 1-700; done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 120 python tools/time_candidates.py tcs 2>&1 | tail -n 1
