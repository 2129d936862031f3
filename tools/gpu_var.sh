mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu --steps 2 --warmup 2"
export CGG_PROFILE=1
( for v in 384_4_2 512_4_2; do echo "== $v"; CGG_LIB=$PWD/tools/var/lib_$v.so timeout 600 $B 2>&1 | grep -v "slice-width\|trace\|decisions" | cut -c1-260 | tail -2;  CGG_LIB=$PWD/tools/var/lib_$v.so timeout 600 $B --no-jet 2>&1 | grep -v "slice-width\|trace\|decisions" | cut -c1-260 | tail -1; done ) > gpurun_out/var.log 2>&1
cat gpurun_out/var.log
