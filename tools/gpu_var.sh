mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_edges.py -x -q 2>&1 | tail -30 ) > gpurun_out/edges.log 2>&1
cat gpurun_out/edges.log
