mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
B="python bench.py --no-e2e --no-cpu --steps 3 --warmup 3"
export CGG_PROFILE=1
echo "== cfg3 p=100"; $B --workload cfg3 --cols 100 2>&1 | tail -2
echo "== cfg3 p=100 theta=1.0"; CGG_COARSE_THETA=1.0 $B --workload cfg3 --cols 100 2>&1 | tail -2
echo "== cfg3 p=100 theta=2.5"; CGG_COARSE_THETA=2.5 $B --workload cfg3 --cols 100 2>&1 | tail -2
echo "== cfg3 p=100 chains=1"; $B --workload cfg3 --cols 100 --chains 1 2>&1 | tail -2
echo "== cfg2"; $B --workload cfg2 2>&1 | tail -2
