# (the variant libraries under tools/ and tools/_old/ were scratch builds -- an older commit, or the working tree with one -D
# flag / one edit -- loaded through CGG_LIB for a same-box A/B; they are not kept: rebuild them the same way to re-run this)
# predicated copies in the lean pair loop: same-box A/B (tools/libcggibbs_pred.so vs the built library)
mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu --steps 3 --warmup 3"
( for v in CUR PRED CUR PRED; do
  case $v in CUR) L=$PWD/mcmcglm_b200/csrc/libcggibbs.so;; PRED) L=$PWD/tools/libcggibbs_pred.so;; esac
  echo "== $v cfg3 p=100"; CGG_LIB=$L timeout 300 $B --workload cfg3 --cols 100 2>&1 | cut -c1-100 | tail -1
done
for v in CUR PRED; do
  case $v in CUR) L=$PWD/mcmcglm_b200/csrc/libcggibbs.so;; PRED) L=$PWD/tools/libcggibbs_pred.so;; esac
  echo "== $v cfg3 full"; CGG_LIB=$L timeout 300 $B 2>&1 | cut -c1-100 | tail -1
done ) > gpurun_out/r2w.log 2>&1
cat gpurun_out/r2w.log
