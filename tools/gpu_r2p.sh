mkdir -p gpurun_out
CGG_NVCC_EXTRA=-DCGG_DECIDER_TICKS python -m mcmcglm_b200.build -f > /dev/null 2>&1
B="python bench.py --no-e2e --no-cpu --steps 2 --warmup 1 --burnin-iters 10"
( for q in 1 0; do
echo "== cfg3 p=100 quad=$q ticks"; CGG_QUAD=$q CGG_PROFILE=1 timeout 300 $B --workload cfg3 --cols 100 2>&1 | grep "cgg profile\] [0-9d]" | tail -3 | cut -c1-330
done ) > gpurun_out/r2p.log 2>&1
python -m mcmcglm_b200.build -f > /dev/null 2>&1
cat gpurun_out/r2p.log
