mkdir -p gpurun_out
unset CGG_PROFILE
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) > gpurun_out/gpu_tests.log 2>&1
cat gpurun_out/gpu_tests.log
timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1; tail -1 gpurun_out/bench_default.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print({k: d[k] for k in ('value', 'ms_per_step', 'e2e', 'roofline', 'clocks')})"
