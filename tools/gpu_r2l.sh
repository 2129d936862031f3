# single-chain streaming at cfg5's per-GPU shape (gaussian n=6.25e6 p=200, 1 chain): ring depth / warps
mkdir -p gpurun_out
B="python bench.py --workload cfg5shard --no-e2e --no-cpu --steps 3 --warmup 3 --burnin-iters 3"
: > gpurun_out/r2l.log
for v in "" "-DCGG_RING_D=8" "-DCGG_THREADS=384" "-DCGG_THREADS=512" "-DCGG_THREADS=384 -DCGG_RING_D=8" "-DCGG_JET_TPI=1" "-DCGG_JET_TPI=1 -DCGG_RING_D=8"; do
  CGG_NVCC_EXTRA="$v" python -m mcmcglm_b200.build -f > /dev/null 2>&1
  echo "== [$v]" >> gpurun_out/r2l.log
  CGG_COLCACHE=0 timeout 300 $B 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('   ', round(d['value']), 'upd/s', round(1e3*d['ms_per_step']/200,1), 'us/pass', 'l2alg', round(d['roofline']['l2_algorithmic_gbs']), 'GB/s')
except Exception as e: print('   fail', e)" >> gpurun_out/r2l.log
done
python -m mcmcglm_b200.build -f > /dev/null 2>&1
cat gpurun_out/r2l.log
