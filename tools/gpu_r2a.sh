# round 2, run A: parity tests of the R log(1-p) form + ownership fix, decision-phase timers, baseline bench
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 ) > gpurun_out/r2a_tests.log 2>&1
cat gpurun_out/r2a_tests.log
B="python bench.py --no-e2e --no-cpu --steps 3 --warmup 3"
( echo "== cfg3 p=100"; timeout 300 $B --workload cfg3 --cols 100 2>&1 | tail -1 | cut -c1-400
  echo "== cfg3 full"; timeout 600 python bench.py --no-cpu --steps 3 --warmup 3 2>&1 | tail -1 ) > gpurun_out/r2a_bench.log 2>&1
# decision phase timers
CGG_NVCC_EXTRA=-DCGG_DECIDER_TICKS python -m mcmcglm_b200.build -f > /dev/null 2>&1
( export CGG_PROFILE=1; echo "== ticks cfg3 p=100"; timeout 300 $B --workload cfg3 --cols 100 2>&1 | grep "cgg profile" | tail -8
  export CGG_PROFILE_TRACE=1; timeout 300 $B --workload cfg3 --cols 100 --steps 1 2>&1 | grep "cgg trace" | tail -16 ) > gpurun_out/r2a_ticks.log 2>&1
python -m mcmcglm_b200.build -f > /dev/null 2>&1
tail -5 gpurun_out/r2a_bench.log | cut -c1-300; tail -30 gpurun_out/r2a_ticks.log
