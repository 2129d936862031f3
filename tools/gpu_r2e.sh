# round 2, run E: 8 warps/SM default, early look-ahead, per-chain w: tests + bench
mkdir -p gpurun_out
( timeout 1800 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 ) > gpurun_out/r2e_tests.log 2>&1
tail -15 gpurun_out/r2e_tests.log
B="python bench.py --no-e2e --no-cpu --steps 3 --warmup 3"
P100="--workload cfg3 --cols 100"
one() { echo "== $1"; shift; timeout 600 "$@" 2>&1 | tail -1 | cut -c1-330; }
( one "cfg3 p=100" $B $P100
  one "cfg3 full (with e2e)" python bench.py --no-cpu --steps 3 --warmup 3
  one "cfg4 p=100" $B --workload cfg4 --cols 100
  one "cfg2" $B --workload cfg2
  one "gauss p=100" $B $P100 --family gaussian ) > gpurun_out/r2e_bench.log 2>&1
( export CGG_PROFILE=1; echo "== profile cfg3 p=100"; timeout 300 $B $P100 2>&1 | grep "cgg profile" | tail -3 ) >> gpurun_out/r2e_bench.log 2>&1
cat gpurun_out/r2e_bench.log | cut -c1-330
grep -o '"e2e": {[^}]*}' gpurun_out/r2e_bench.log | head -2
