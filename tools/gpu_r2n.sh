# group passes (four chains per walk): parity tests, then A/B benches
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_edges.py -x -q -k "group_passes or handover_stress or pair_passes" 2>&1 | tail -15 ) > gpurun_out/r2n_tests.log 2>&1
cat gpurun_out/r2n_tests.log
B="python bench.py --no-e2e --no-cpu --steps 3 --warmup 3"
( for q in 0 1; do
  echo "== cfg3 p=100 quad=$q"; CGG_QUAD=$q CGG_PROFILE=1 timeout 300 $B --workload cfg3 --cols 100 2>&1 | grep -v "slice-width" | cut -c1-400 | tail -4
  echo "== gauss p=100 quad=$q"; CGG_QUAD=$q timeout 300 $B --workload cfg3 --cols 100 --family gaussian 2>&1 | cut -c1-300 | tail -1
done
echo "== cfg3 full quad=1"; timeout 600 $B 2>&1 | cut -c1-700 | tail -1 ) > gpurun_out/r2n_bench.log 2>&1
cat gpurun_out/r2n_bench.log
