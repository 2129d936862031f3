mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_jet.py tests/test_gpu_edges.py tests/test_gpu_parity.py tests/test_gpu_api.py tests/test_gpu_sharded.py -x -q 2>&1 | tail -4 ) > gpurun_out/pair_tests.log 2>&1
cat gpurun_out/pair_tests.log
B="python bench.py --no-e2e --no-cpu --steps 3 --warmup 2"
export CGG_PROFILE=1
( echo "== cfg3"; timeout 300 $B 2>&1 | grep "ms; per\|value\|decisions" | cut -c1-250 | tail -3
echo "== cfg4"; timeout 300 $B --workload cfg4 2>&1 | grep "value" | cut -c1-140 | tail -1
echo "== cfg2"; timeout 300 $B --workload cfg2 2>&1 | grep "value" | cut -c1-140 | tail -1
echo "== C=1"; timeout 300 $B --cols 100 --chains 1 2>&1 | grep "value" | cut -c1-140 | tail -1 ) > gpurun_out/var.log 2>&1
cat gpurun_out/var.log
