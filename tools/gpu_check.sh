# (the CGG_PROFILE lines below need a library built with CGG_NVCC_EXTRA=-DCGG_PROFILE_BUILD python -m mcmcglm_b200.build -f)
# quick GPU regression: parity/jet/api tests, then short benches with the phase counters
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_jet.py tests/test_gpu_parity.py tests/test_gpu_api.py tests/test_gpu_coarse.py -x -q 2>&1 | tail -4 ) > gpurun_out/gpu_tests.log 2>&1
cat gpurun_out/gpu_tests.log
B="python bench.py --no-e2e --no-cpu --steps 3 --warmup 2"
export CGG_PROFILE=1
( echo "== cfg3 p=100"; timeout 300 $B --workload cfg3 --cols 100 2>&1 | grep -v "slice-width" | cut -c1-300 | tail -3
echo "== cfg3 full"; timeout 600 $B --steps 2 --warmup 1 2>&1 | grep -v "slice-width\|trace" | cut -c1-300 | tail -3
echo "== gauss p=100"; timeout 300 $B --workload cfg3 --cols 100 --family gaussian 2>&1 | grep -v "slice-width" | cut -c1-300| tail -3
echo "== cfg4 p=100"; timeout 300 $B --workload cfg4 --cols 100 2>&1 | grep -v "slice-width" | cut -c1-300| tail -3
echo "== cfg2"; timeout 300 $B --workload cfg2 2>&1 | grep -v "slice-width" | cut -c1-300| tail -3
echo "== cfg3 p=100 C=1"; timeout 300 $B --workload cfg3 --cols 100 --chains 1 2>&1 | grep -v "slice-width\|trace" | cut -c1-300| tail -3 ) > gpurun_out/var.log 2>&1
cat gpurun_out/var.log
( echo "== cfg3 full no-jet"; timeout 600 $B --steps 2 --warmup 1 --no-jet 2>&1 | grep -v "slice-width\|trace" | cut -c1-300 | tail -3 ) >> gpurun_out/var.log 2>&1
tail -4 gpurun_out/var.log
