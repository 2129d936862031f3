mkdir -p gpurun_out
( timeout 1800 python -m pytest tests/test_gpu_coarse.py tests/test_gpu_parity.py tests/test_gpu_jet.py tests/test_gpu_edges.py tests/test_gpu_wider.py -q -m gpu 2>&1 | tail -8 ) > gpurun_out/r2k_tests.log 2>&1
tail -8 gpurun_out/r2k_tests.log
python tools/transient_probe.py --iters 14 > gpurun_out/r2k_transient.log 2>&1
cut -c1-330 gpurun_out/r2k_transient.log
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2k_bench_default.log 2>&1; tail -1 gpurun_out/r2k_bench_default.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value']), 'e2e', d['e2e'])"
