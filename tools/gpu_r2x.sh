# (the variant libraries under tools/ and tools/_old/ were scratch builds -- an older commit, or the working tree with one -D
# flag / one edit -- loaded through CGG_LIB for a same-box A/B; they are not kept: rebuild them the same way to re-run this)
# 32-bit tile counters in the lean loop (tools/libcggibbs_cnt.so) vs the built library, parity tests on the variant, and a
# --set full capture of the built library's steady launch
mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu --steps 3 --warmup 3"
( for v in CUR CNT CUR CNT; do
  case $v in CUR) L=$PWD/mcmcglm_b200/csrc/libcggibbs.so;; CNT) L=$PWD/tools/libcggibbs_cnt.so;; esac
  echo "== $v cfg3 p=100"; CGG_LIB=$L timeout 300 $B --workload cfg3 --cols 100 2>&1 | cut -c1-100 | tail -1
done
for v in CUR CNT; do
  case $v in CUR) L=$PWD/mcmcglm_b200/csrc/libcggibbs.so;; CNT) L=$PWD/tools/libcggibbs_cnt.so;; esac
  echo "== $v cfg3 full"; CGG_LIB=$L timeout 300 $B 2>&1 | cut -c1-100 | tail -1
done ) > gpurun_out/r2x.log 2>&1
cat gpurun_out/r2x.log
( CGG_LIB=$PWD/tools/libcggibbs_cnt.so timeout 600 python -m pytest tests/test_gpu_edges.py tests/test_gpu_parity.py -x -q 2>&1 | tail -3 ) > gpurun_out/r2x_tests.log 2>&1
cat gpurun_out/r2x_tests.log
P100="python bench.py --workload cfg3 --cols 100 --steps 1 --warmup 3 --burnin-iters 30 --no-e2e --no-cpu"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_persistent -s 4 -c 1 -f -o gpurun_out/r02d_sweep_binomial $P100 > gpurun_out/r02d_ncu_full.log 2>&1
ls -la gpurun_out/r02d_sweep_binomial.ncu-rep
