mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cuda_graph or self_transfer or host_pipeline or install or second_pass or gather_fold or up2" > gpurun_out/r2_pytest_m.log 2>&1; echo "pytest exit $?"; tail -n 6 gpurun_out/r2_pytest_m.log
timeout 300 python tools/time_module.py > gpurun_out/r2_module_latency.json 2> gpurun_out/r2_module_latency.err; echo "module exit $?"; cat gpurun_out/r2_module_latency.json; tail -n 3 gpurun_out/r2_module_latency.err
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_sanitizer_memcheck.log 2>&1; echo "memcheck exit $?"; tail -n 12 gpurun_out/r2_sanitizer_memcheck.log
