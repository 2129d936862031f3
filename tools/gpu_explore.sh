mkdir -p gpurun_out
./tools/fp64_probe
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
B="python bench.py --no-e2e --no-cpu --steps 3 --warmup 3"
export CGG_PROFILE=1
echo "== cfg3 p=100"; $B --workload cfg3 --p 100 2>&1 | tail -2
echo "== cfg3 p=100 chains=1"; $B --workload cfg3 --p 100 --chains 1 2>&1 | tail -2
echo "== cfg3 p=100 tau=0.8"; $B --workload cfg3 --p 100 --tau 0.8 2>&1 | tail -2
echo "== cfg3 p=100 tau=0.2"; $B --workload cfg3 --p 100 --tau 0.2 2>&1 | tail -2
echo "== poisson cfg4 p=100"; $B --workload cfg4 --p 100 2>&1 | tail -2
echo "== cfg2"; $B --workload cfg2 2>&1 | tail -2
