mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py -m gpu -x -q -k "bf16 or fusion or gather_fold or host_pipeline or integration or golden" > gpurun_out/r2_pytest_b.log 2>&1; echo "pytest exit $?"; tail -n 25 gpurun_out/r2_pytest_b.log
