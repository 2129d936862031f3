mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_jet.py tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_gpu_wider.py -x -q 2>&1 | tail -12 ) > gpurun_out/r2r_tests.log 2>&1
cat gpurun_out/r2r_tests.log
B="python bench.py --no-e2e --no-cpu --steps 3 --warmup 3"
( echo "== cfg4 p=100 NEW"; timeout 300 $B --workload cfg4 --cols 100 2>&1 | cut -c1-120 | tail -1
echo "== cfg4 p=100 NEW profile"; CGG_PROFILE=1 timeout 300 $B --workload cfg4 --cols 100 2>&1 | grep "cgg profile\] [0-9d]" | tail -3 | head -2 | cut -c1-300
echo "== cfg4 p=100 OLD"; CGG_LIB=$PWD/tools/_old/libcggibbs_old.so timeout 300 $B --workload cfg4 --cols 100 2>&1 | cut -c1-120 | tail -1
echo "== cfg4 full NEW"; timeout 600 $B --workload cfg4 2>&1 | cut -c1-1500 | tail -1
) > gpurun_out/r2r_bench.log 2>&1
cat gpurun_out/r2r_bench.log
