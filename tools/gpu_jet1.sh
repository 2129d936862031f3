mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_jet.py -x -q 2>&1 | tail -15 > gpurun_out/jet_tests.log
cat gpurun_out/jet_tests.log
B="python bench.py --no-e2e --no-cpu --steps 3 --warmup 3"
export CGG_PROFILE=1
( echo "== cfg3 p=100 jet"; timeout 300 $B --workload cfg3 --cols 100 2>&1 | tail -3
echo "== cfg3 p=100 nojet"; timeout 300 $B --workload cfg3 --cols 100 --no-jet 2>&1 | tail -3
echo "== gauss p=100 jet"; timeout 300 $B --workload cfg3 --cols 100 --family gaussian 2>&1 | tail -3
echo "== poisson p=100 jet"; timeout 300 $B --workload cfg4 --cols 100 2>&1 | tail -3
echo "== cfg3 p=100 chains=1 jet"; timeout 300 $B --workload cfg3 --cols 100 --chains 1 2>&1 | tail -3
echo "== cfg2 jet"; timeout 300 $B --workload cfg2 2>&1 | tail -3 ) > gpurun_out/jet_bench1.log 2>&1
cat gpurun_out/jet_bench1.log | cut -c1-1500
