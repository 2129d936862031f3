mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu --steps 2 --warmup 2"
export CGG_PROFILE=1
( for v in 256_4_1 256_4_2 256_8_2 512_4_2; do echo "== $v"; CGG_LIB=$PWD/tools/var/lib_$v.so timeout 600 python -m pytest tests/test_gpu_jet.py -x -q 2>&1 | tail -1; CGG_LIB=$PWD/tools/var/lib_$v.so timeout 600 $B 2>&1 | grep -v "slice-width\|trace" | cut -c1-300 | tail -3; done ) > gpurun_out/var.log 2>&1
cat gpurun_out/var.log
