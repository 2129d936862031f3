mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_jet.py tests/test_gpu_edges.py tests/test_gpu_parity.py tests/test_gpu_sharded.py -x -q -k "poisson or enclosure or chunk or more_chains" 2>&1 | tail -3 ) > gpurun_out/pois_tests.log 2>&1
cat gpurun_out/pois_tests.log
B="python bench.py --no-cpu --no-e2e --steps 3 --warmup 2 --workload cfg4"
for i in 1 2; do timeout 300 $B 2>&1 | tail -1 | cut -c1-100; done
