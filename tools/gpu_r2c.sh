# round 2, run C: decision pre-phase + combined verdict round, L2-persisting eta; tests, bench A/B, ncu, ring/warp variants
mkdir -p gpurun_out
( timeout 1800 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 ) > gpurun_out/r2c_tests.log 2>&1
tail -15 gpurun_out/r2c_tests.log
B="python bench.py --no-e2e --no-cpu --steps 3 --warmup 3"
P100="--workload cfg3 --cols 100"
one() { echo "== $1"; shift; timeout 600 "$@" 2>&1 | tail -1 | cut -c1-330; }
( one "cfg3 p=100" $B $P100
  CGG_L2_PERSIST=0 one "cfg3 p=100 no-L2-persist" $B $P100
  one "cfg3 full" python bench.py --no-cpu --steps 3 --warmup 3
  one "cfg4 p=100" $B --workload cfg4 --cols 100
  one "cfg2" $B --workload cfg2
  one "gauss p=100" $B $P100 --family gaussian ) > gpurun_out/r2c_bench.log 2>&1
cat gpurun_out/r2c_bench.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep_persistent -s 4 -c 1 -f -o gpurun_out/prof_r2c_binom $B $P100 --steps 1 > gpurun_out/r2c_ncu.log 2>&1
for v in "-DCGG_THREADS=256 -DCGG_RING_D=8" "-DCGG_THREADS=512 -DCGG_RING_D=4" "-DCGG_THREADS=256 -DCGG_RING_D=4"; do
  CGG_NVCC_EXTRA="$v" python -m mcmcglm_b200.build -f > /dev/null 2>&1
  ( one "cfg3 p=100 [$v]" $B $P100 ) >> gpurun_out/r2c_bench.log 2>&1
done
CGG_NVCC_EXTRA=-DCGG_DECIDER_TICKS python -m mcmcglm_b200.build -f > /dev/null 2>&1
( export CGG_PROFILE=1; echo "== ticks cfg3 p=100"; timeout 300 $B $P100 2>&1 | grep "cgg profile" | tail -4
  export CGG_PROFILE_TRACE=1; timeout 300 $B $P100 --steps 1 2>&1 | grep "cgg trace" | tail -8 ) > gpurun_out/r2c_ticks.log 2>&1
python -m mcmcglm_b200.build -f > /dev/null 2>&1
tail -4 gpurun_out/r2c_bench.log; cat gpurun_out/r2c_ticks.log | cut -c1-300
