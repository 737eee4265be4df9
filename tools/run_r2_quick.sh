mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "search or tile or edge or certified or second_pass or library or module" > gpurun_out/r2_pytest_q.log 2>&1; echo "pytest exit $?"; tail -n 6 gpurun_out/r2_pytest_q.log
timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_q.json 2> gpurun_out/r2_bench_q.err; echo "bench exit $?"; tail -n 3 gpurun_out/r2_bench_q.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_q.json') if l.startswith('{')][0])
r=d['roofline']
print('ms/step',d['ms_per_step'],'kernel_ms',r['kernel_ms'],'frac',r['frac'],'clk',d['clocks'].get('search_kernel_sm_mhz'),'plan',d['plan'])
print('stats',d['search_stats_last_step'])
print('fixed',d['window_ab'])
PY
