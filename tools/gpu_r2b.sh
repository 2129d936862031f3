# round 2, run B: limb accumulators + lighter light pass: tests, bench, decision-phase timers
mkdir -p gpurun_out
( timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -25 ) > gpurun_out/r2b_tests.log 2>&1
tail -25 gpurun_out/r2b_tests.log
B="python bench.py --no-e2e --no-cpu --steps 3 --warmup 3"
( echo "== cfg3 p=100"; timeout 300 $B --workload cfg3 --cols 100 2>&1 | tail -1 | cut -c1-400
  echo "== cfg3 full"; timeout 600 python bench.py --no-cpu --steps 3 --warmup 3 2>&1 | tail -1
  echo "== cfg4 p=100"; timeout 300 $B --workload cfg4 --cols 100 2>&1 | tail -1 | cut -c1-400
  echo "== cfg2"; timeout 300 $B --workload cfg2 2>&1 | tail -1 | cut -c1-400
  echo "== gauss p=100"; timeout 300 $B --workload cfg3 --family gaussian --cols 100 2>&1 | tail -1 | cut -c1-400 ) > gpurun_out/r2b_bench.log 2>&1
CGG_NVCC_EXTRA=-DCGG_PAIR_TPI=2 python -m mcmcglm_b200.build -f > /dev/null 2>&1
( echo "== cfg3 p=100 PAIR_TPI=2"; timeout 300 $B --workload cfg3 --cols 100 2>&1 | tail -1 | cut -c1-400 ) >> gpurun_out/r2b_bench.log 2>&1
CGG_NVCC_EXTRA=-DCGG_DECIDER_TICKS python -m mcmcglm_b200.build -f > /dev/null 2>&1
( export CGG_PROFILE=1; echo "== ticks cfg3 p=100"; timeout 300 $B --workload cfg3 --cols 100 2>&1 | grep "cgg profile" | tail -4
  export CGG_PROFILE_TRACE=1; timeout 300 $B --workload cfg3 --cols 100 --steps 1 2>&1 | grep "cgg trace" | tail -16 ) > gpurun_out/r2b_ticks.log 2>&1
python -m mcmcglm_b200.build -f > /dev/null 2>&1
cut -c1-330 gpurun_out/r2b_bench.log; cat gpurun_out/r2b_ticks.log
