# 2-GPU checks: the NCCL row-sharded test, chain-parallel weak scaling of the headline workload, the row-sharded workload
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/multi_smi.txt
( timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q 2>&1 | tail -3 ) > gpurun_out/multi_tests.log 2>&1
cat gpurun_out/multi_tests.log
python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/scale_n1.log 2>&1; tail -1 gpurun_out/scale_n1.log | cut -c1-160
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/scale_n2.log 2>&1; tail -1 gpurun_out/scale_n2.log | cut -c1-160
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload cfg5 --rows 12500000 --steps 2 --warmup 1 --burnin-iters 3 --no-cpu --no-e2e > gpurun_out/cfg5_n2.log 2>&1; tail -1 gpurun_out/cfg5_n2.log | cut -c1-700
