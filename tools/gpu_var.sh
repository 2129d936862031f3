mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_jet.py -x -q 2>&1 | tail -5
echo "--- D=5"; CGG_LIB=$PWD/tools/var/lib_d5.so timeout 900 python -m pytest tests/test_gpu_jet.py -x -q 2>&1 | tail -5 ) > gpurun_out/jet_tests.log 2>&1
cat gpurun_out/jet_tests.log
B="python bench.py --no-e2e --no-cpu --steps 3 --warmup 2 --workload cfg3 --cols 100"
export CGG_PROFILE=1
( echo "== D7"; timeout 300 $B 2>&1 | grep -v "slice-width" | cut -c1-330 | tail -3
echo "== D5"; CGG_LIB=$PWD/tools/var/lib_d5.so timeout 300 $B 2>&1 | grep -v "slice-width" | cut -c1-330 | tail -3
echo "== D5 full cfg3"; CGG_LIB=$PWD/tools/var/lib_d5.so timeout 600 python bench.py --no-e2e --no-cpu --steps 2 --warmup 1 2>&1 | grep -v "slice-width" | tail -3 ) > gpurun_out/var.log 2>&1
cat gpurun_out/var.log
