mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_edges.py -x -q -k "pair_passes" 2>&1 | tail -12 ) > gpurun_out/edges.log 2>&1
cat gpurun_out/edges.log
