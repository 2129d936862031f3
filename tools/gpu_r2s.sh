# same-box A/B of the plain-update path (CGG_EARLY) on the latency-bound workloads.  tools/_old/libcggibbs_old.so was a build of the
# round's first commit (git show 8b5f41b:... into a scratch tree, ABI constant patched to 3), loaded through CGG_LIB; not kept.
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_jet.py tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_gpu_api.py tests/test_gpu_wider.py -x -q 2>&1 | tail -12 ) > gpurun_out/r2s_tests.log 2>&1
cat gpurun_out/r2s_tests.log
B="python bench.py --no-e2e --no-cpu --steps 3 --warmup 3"
( for e in 1 0; do
echo "== cfg2 (cluster) early=$e"; CGG_EARLY=$e timeout 300 $B --workload cfg2 2>&1 | cut -c1-120 | tail -1
echo "== cfg3 p=100 early=$e"; CGG_EARLY=$e timeout 300 $B --workload cfg3 --cols 100 2>&1 | cut -c1-120 | tail -1
echo "== cfg3 p=100 one chain early=$e"; CGG_EARLY=$e timeout 300 $B --workload cfg3 --cols 100 --chains 1 2>&1 | cut -c1-120 | tail -1
echo "== cfg5shard early=$e"; CGG_EARLY=$e timeout 300 $B --workload cfg5shard --cols 50 2>&1 | cut -c1-120 | tail -1
echo "== cfg4 p=100 early=$e"; CGG_EARLY=$e timeout 300 $B --workload cfg4 --cols 100 2>&1 | cut -c1-120 | tail -1
echo "== tiny early=$e"; CGG_EARLY=$e timeout 300 $B --workload tiny 2>&1 | cut -c1-120 | tail -1
done
echo "== cfg2 profile"; CGG_PROFILE=1 timeout 300 $B --workload cfg2 2>&1 | grep "cgg profile\] cluster" | tail -1 | cut -c1-400
echo "== OLD cfg3 p=100"; CGG_LIB=$PWD/tools/_old/libcggibbs_old.so timeout 300 $B --workload cfg3 --cols 100 2>&1 | cut -c1-120 | tail -1
echo "== OLD cfg2"; CGG_LIB=$PWD/tools/_old/libcggibbs_old.so timeout 300 $B --workload cfg2 2>&1 | cut -c1-120 | tail -1
) > gpurun_out/r2s_bench.log 2>&1
cat gpurun_out/r2s_bench.log
