mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_jet.py tests/test_gpu_parity.py -x -q 2>&1 | tail -12 ) > gpurun_out/gpu_tests.log 2>&1
cat gpurun_out/gpu_tests.log
