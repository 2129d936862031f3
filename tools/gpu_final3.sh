# round-end sequence + the other workloads, one GPU
bash tools/gpu_final.sh
B="python bench.py --no-e2e --no-cpu --steps 3 --warmup 3"
( echo "== cfg4 full"; timeout 300 $B --workload cfg4 2>&1 | cut -c1-100 | tail -1
echo "== gauss p=100"; timeout 300 $B --workload cfg3 --cols 100 --family gaussian 2>&1 | cut -c1-100 | tail -1
echo "== cfg2"; timeout 300 $B --workload cfg2 2>&1 | cut -c1-100 | tail -1
echo "== one chain p=100"; timeout 300 $B --workload cfg3 --cols 100 --chains 1 2>&1 | cut -c1-100 | tail -1
echo "== cfg5shard"; timeout 300 $B --workload cfg5shard --cols 50 2>&1 | cut -c1-100 | tail -1 ) > gpurun_out/final_others.log 2>&1
cat gpurun_out/final_others.log
