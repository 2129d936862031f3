# multi-GPU checks (gpurun --gpus N): the 2-process sharded tests, then the bench contract under torchrun with cfg5 riding along
N=${N:-2}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_sharded.py -q -m gpu 2>&1 | tail -6 ) > gpurun_out/multi_tests_n$N.log 2>&1
cat gpurun_out/multi_tests_n$N.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu > gpurun_out/multi_bench_n$N.log 2> gpurun_out/multi_bench_n$N.err
tail -1 gpurun_out/multi_bench_n$N.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value', round(d['value']), 'n_gpus', d['n_gpus'], 'e2e', d['e2e'] and round(d['e2e']['value']))
print('extra', json.dumps(d.get('extra_workloads'))[:1500])"
tail -5 gpurun_out/multi_bench_n$N.err | cut -c1-300
