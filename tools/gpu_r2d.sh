# round 2, run D: warps-per-SM / ring depth / tiles-per-iteration sweep of the sweep kernel (cfg3 p=100, steady state)
mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu --steps 3 --warmup 3 --workload cfg3 --cols 100"
: > gpurun_out/r2d_bench.log
for v in "-DCGG_THREADS=256 -DCGG_RING_D=4 -DCGG_PAIR_TPI=2" "-DCGG_THREADS=192 -DCGG_RING_D=4" "-DCGG_THREADS=128 -DCGG_RING_D=8 -DCGG_PAIR_TPI=2" \
         "-DCGG_THREADS=128 -DCGG_RING_D=4" "-DCGG_THREADS=256 -DCGG_RING_D=8 -DCGG_PAIR_TPI=2" "-DCGG_THREADS=320 -DCGG_RING_D=4" "-DCGG_THREADS=256 -DCGG_RING_D=4"; do
  CGG_NVCC_EXTRA="$v" python -m mcmcglm_b200.build -f > /dev/null 2>&1
  echo "== [$v]" >> gpurun_out/r2d_bench.log
  timeout 300 $B 2>&1 | tail -1 | cut -c1-120 >> gpurun_out/r2d_bench.log
done
# per-CTA balance of the default build
python -m mcmcglm_b200.build -f > /dev/null 2>&1
( export CGG_PROFILE=1 CGG_PROFILE_CTAS=1; timeout 300 $B --steps 1 2>&1 | grep "cgg cta" | sort -k9 -n | awk 'NR%8==1' ) > gpurun_out/r2d_ctas.log 2>&1
cat gpurun_out/r2d_bench.log; tail -20 gpurun_out/r2d_ctas.log
