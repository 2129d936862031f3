mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
B="python bench.py --no-e2e --no-cpu --steps 3 --warmup 3"
export CGG_PROFILE=1
echo "== cfg3 p=100"; $B --workload cfg3 --cols 100 2>&1 | tail -2
echo "== cfg3 p=100 chains=4"; $B --workload cfg3 --cols 100 --chains 4 2>&1 | tail -2
echo "== cfg3 p=100 chains=16"; $B --workload cfg3 --cols 100 --chains 16 2>&1 | tail -2
echo "== gaussian n=1e6 p=100 C=8"; $B --workload cfg3 --cols 100 --family gaussian 2>&1 | tail -2
echo "== poisson cfg4 p=100"; $B --workload cfg4 --cols 100 2>&1 | tail -2
echo "== cfg2"; $B --workload cfg2 2>&1 | tail -2
