mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_final.log 2>&1; echo "pytest exit $?"; tail -n 5 gpurun_out/r2_pytest_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 gpurun_out/r2_smoke.log
timeout 900 python bench.py > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "bench exit $?"; tail -n 5 gpurun_out/r2_bench1.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref exit $?"; tail -n 3 gpurun_out/r2_bench_ref.err; cut -c1-600 gpurun_out/r2_bench_ref.json
