mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu --steps 3 --warmup 3"
( for v in OLD CUR vC CUR OLD; do
  case $v in OLD) L=$PWD/tools/_old/libcggibbs_old.so;; CUR) L=$PWD/mcmcglm_b200/csrc/libcggibbs.so;; *) L=$PWD/tools/libcggibbs_$v.so;; esac
  echo "== $v cfg3 p=100"; CGG_LIB=$L timeout 300 $B --workload cfg3 --cols 100 2>&1 | cut -c1-100 | tail -1
done
for v in OLD CUR vC; do
  case $v in OLD) L=$PWD/tools/_old/libcggibbs_old.so;; CUR) L=$PWD/mcmcglm_b200/csrc/libcggibbs.so;; *) L=$PWD/tools/libcggibbs_$v.so;; esac
  echo "== $v cfg3 full"; CGG_LIB=$L timeout 300 $B 2>&1 | cut -c1-100 | tail -1
done
echo "== CUR cfg2"; timeout 300 $B --workload cfg2 2>&1 | cut -c1-100 | tail -1
echo "== CUR cfg4 p=100"; timeout 300 $B --workload cfg4 --cols 100 2>&1 | cut -c1-100 | tail -1
echo "== OLD cfg4 p=100"; CGG_LIB=$PWD/tools/_old/libcggibbs_old.so timeout 300 $B --workload cfg4 --cols 100 2>&1 | cut -c1-100 | tail -1
echo "== CUR gauss p=100"; timeout 300 $B --workload cfg3 --cols 100 --family gaussian 2>&1 | cut -c1-100 | tail -1
echo "== OLD gauss p=100"; CGG_LIB=$PWD/tools/_old/libcggibbs_old.so timeout 300 $B --workload cfg3 --cols 100 --family gaussian 2>&1 | cut -c1-100 | tail -1
) > gpurun_out/r2u.log 2>&1
cat gpurun_out/r2u.log
