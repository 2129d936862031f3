mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu --steps 3 --warmup 3"
( echo "== cfg2 (cluster)"; timeout 300 $B --workload cfg2 2>&1 | cut -c1-120 | tail -1
echo "== cfg2 profile"; CGG_PROFILE=1 timeout 300 $B --workload cfg2 2>&1 | grep "cgg profile\] cluster" | tail -1 | cut -c1-400
echo "== cfg3 p=100 default"; timeout 300 $B --workload cfg3 --cols 100 2>&1 | cut -c1-120 | tail -1
echo "== cfg3 p=100 one chain"; timeout 300 $B --workload cfg3 --cols 100 --chains 1 2>&1 | cut -c1-120 | tail -1
echo "== tiny"; timeout 300 $B --workload tiny 2>&1 | cut -c1-120 | tail -1
echo "== cfg5shard"; timeout 300 $B --workload cfg5shard --cols 50 2>&1 | cut -c1-120 | tail -1
) > gpurun_out/r2t_bench.log 2>&1
cat gpurun_out/r2t_bench.log
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py -x -q 2>&1 | tail -3 ) > gpurun_out/r2t_tests.log 2>&1
cat gpurun_out/r2t_tests.log
