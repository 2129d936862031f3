mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu --steps 3 --warmup 3"
( echo "== OLD lib cfg3 p=100"; CGG_LIB=$PWD/tools/_old/libcggibbs_old.so timeout 300 $B --workload cfg3 --cols 100 2>&1 | cut -c1-120 | tail -1
for e in 1 0; do echo "== NEW cfg3 p=100 early=$e"; CGG_EARLY=$e timeout 300 $B --workload cfg3 --cols 100 2>&1 | cut -c1-120 | tail -1; done
echo "== NEW early=0 profile"; CGG_EARLY=0 CGG_PROFILE=1 timeout 300 $B --workload cfg3 --cols 100 2>&1 | grep "cgg profile\] [0-9d]" | tail -3 | head -2 | cut -c1-300
echo "== OLD lib cfg3 p=100"; CGG_LIB=$PWD/tools/_old/libcggibbs_old.so timeout 300 $B --workload cfg3 --cols 100 2>&1 | cut -c1-120 | tail -1
) > gpurun_out/r2q.log 2>&1
cat gpurun_out/r2q.log
