mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu --steps 3 --warmup 2"
( for i in 1 2; do
echo "== tpi1"; timeout 300 $B 2>&1 | tail -1 | cut -c1-140
echo "== tpi2"; CGG_LIB=$PWD/tools/var/lib_ptpi2.so timeout 300 $B 2>&1 | tail -1 | cut -c1-140
done
echo "== tpi2 tests"; CGG_LIB=$PWD/tools/var/lib_ptpi2.so CGG_PAIR=1 timeout 600 python -m pytest tests/test_gpu_jet.py tests/test_gpu_edges.py -x -q 2>&1 | tail -1 ) > gpurun_out/var.log 2>&1
cat gpurun_out/var.log
