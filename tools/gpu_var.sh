mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu --steps 2 --warmup 1"
export CGG_PROFILE=1
( timeout 600 python -m pytest tests/test_gpu_jet.py -x -q 2>&1 | tail -2
echo "== nofence full"; timeout 600 $B 2>&1 | grep -v "slice-width\|trace" | cut -c1-300 | tail -3
echo "== fence.acq_rel.cta full"; CGG_LIB=$PWD/tools/var/lib_fence.so timeout 600 $B 2>&1 | grep -v "slice-width\|trace" | cut -c1-300 | tail -3 ) > gpurun_out/var.log 2>&1
cat gpurun_out/var.log
