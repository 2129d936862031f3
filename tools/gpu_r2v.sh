# (the variant libraries under tools/ and tools/_old/ were scratch builds -- an older commit, or the working tree with one -D
# flag / one edit -- loaded through CGG_LIB for a same-box A/B; they are not kept: rebuild them the same way to re-run this)
# lean pair passes (-DCGG_LEAN_PAIR, tools/libcggibbs_lean.so): same-box A/B and parity
mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu --steps 3 --warmup 3"
( for v in OLD CUR LEAN CUR LEAN; do
  case $v in OLD) L=$PWD/tools/_old/libcggibbs_old.so;; CUR) L=$PWD/mcmcglm_b200/csrc/libcggibbs.so;; LEAN) L=$PWD/tools/libcggibbs_lean.so;; esac
  echo "== $v cfg3 p=100"; CGG_LIB=$L timeout 300 $B --workload cfg3 --cols 100 2>&1 | cut -c1-100 | tail -1
done
for v in CUR LEAN; do
  case $v in CUR) L=$PWD/mcmcglm_b200/csrc/libcggibbs.so;; LEAN) L=$PWD/tools/libcggibbs_lean.so;; esac
  echo "== $v cfg3 full"; CGG_LIB=$L timeout 300 $B 2>&1 | cut -c1-100 | tail -1
  echo "== $v cfg4 p=100"; CGG_LIB=$L timeout 300 $B --workload cfg4 --cols 100 2>&1 | cut -c1-100 | tail -1
  echo "== $v gauss p=100"; CGG_LIB=$L timeout 300 $B --workload cfg3 --cols 100 --family gaussian 2>&1 | cut -c1-100 | tail -1
done ) > gpurun_out/r2v.log 2>&1
cat gpurun_out/r2v.log
( CGG_LIB=$PWD/tools/libcggibbs_lean.so timeout 900 python -m pytest tests/test_gpu_edges.py tests/test_gpu_parity.py tests/test_gpu_jet.py -x -q 2>&1 | tail -4 ) > gpurun_out/r2v_tests.log 2>&1
cat gpurun_out/r2v_tests.log
