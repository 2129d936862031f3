# 8-GPU check of the bench contract: chain-parallel cfg3 + the row-sharded cfg5 riding along
N=${N:-8}
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/multi_bench_n$N.log 2> gpurun_out/multi_bench_n$N.err
tail -1 gpurun_out/multi_bench_n$N.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value', round(d['value']), 'n_gpus', d['n_gpus'], 'ms/step', d['ms_per_step'])
print('extra', json.dumps(d.get('extra_workloads'))[:1800])"
tail -3 gpurun_out/multi_bench_n$N.err | cut -c1-300
