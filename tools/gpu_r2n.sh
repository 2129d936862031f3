# group passes need a library built with -DCGG_GROUP_PASSES (CGG_NVCC_EXTRA=-DCGG_GROUP_PASSES python -m mcmcglm_b200.build -f);
# CGG_PROFILE lines need -DCGG_PROFILE_BUILD as well.
# early publication + group passes: tests, then A/B benches
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_edges.py tests/test_gpu_jet.py tests/test_gpu_parity.py -x -q 2>&1 | tail -5 ) > gpurun_out/r2n_tests.log 2>&1
cat gpurun_out/r2n_tests.log
B="python bench.py --no-e2e --no-cpu --steps 3 --warmup 3"
( for e in 0 1; do for q in 0 1; do
  echo "== cfg3 p=100 early=$e quad=$q"; CGG_EARLY=$e CGG_QUAD=$q timeout 300 $B --workload cfg3 --cols 100 2>&1 | cut -c1-120 | tail -1
done; done
echo "== cfg3 p=100 early=1 quad=1 profile"; CGG_PROFILE=1 timeout 300 $B --workload cfg3 --cols 100 2>&1 | grep "cgg profile\] [0-9d]" | tail -2 | cut -c1-300
for q in 0 1; do echo "== gauss p=100 early=1 quad=$q"; CGG_QUAD=$q timeout 300 $B --workload cfg3 --cols 100 --family gaussian 2>&1 | cut -c1-120 | tail -1; done
echo "== cfg4 p=100"; timeout 300 $B --workload cfg4 --cols 100 2>&1 | cut -c1-120 | tail -1
echo "== cfg4 p=100 early=0"; CGG_EARLY=0 timeout 300 $B --workload cfg4 --cols 100 2>&1 | cut -c1-120 | tail -1
echo "== cfg3 full"; timeout 600 $B 2>&1 | cut -c1-120 | tail -1
echo "== cfg3 full early=0 quad=0"; CGG_EARLY=0 CGG_QUAD=0 timeout 600 $B 2>&1 | cut -c1-120 | tail -1 ) > gpurun_out/r2n_bench.log 2>&1
cat gpurun_out/r2n_bench.log
