mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu --steps 3 --warmup 3 --workload cfg2"
CGG_NVCC_EXTRA=-DCGG_DECIDER_TICKS python -m mcmcglm_b200.build -f > /dev/null 2>&1
( CGG_PROFILE=1 timeout 300 $B 2>&1 | grep "cgg profile" | tail -2 ) > gpurun_out/r2m.log 2>&1
for v in "-DCGG_THREADS=512" "-DCGG_THREADS=384"; do
  CGG_NVCC_EXTRA="$v" python -m mcmcglm_b200.build -f > /dev/null 2>&1
  echo "== [$v]" >> gpurun_out/r2m.log
  ( CGG_PROFILE=1 timeout 300 $B 2>&1 | grep "cgg profile\|metric" | tail -2 | cut -c1-420 ) >> gpurun_out/r2m.log 2>&1
done
python -m mcmcglm_b200.build -f > /dev/null 2>&1
cat gpurun_out/r2m.log
