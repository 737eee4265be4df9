mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1; echo "pytest exit $?"; tail -n 25 gpurun_out/r2_pytest1.log
timeout 600 python tools/time_window.py > gpurun_out/r2_window.json 2> gpurun_out/r2_window.err; echo "window exit $?"; tail -n 5 gpurun_out/r2_window.err
timeout 900 python bench.py > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "bench exit $?"; tail -n 5 gpurun_out/r2_bench1.err
