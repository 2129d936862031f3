mkdir -p gpurun_out
( CGG_PAIR=1 timeout 1200 python -m pytest tests/test_gpu_jet.py tests/test_gpu_edges.py tests/test_gpu_parity.py tests/test_gpu_api.py -x -q 2>&1 | tail -2 ) > gpurun_out/pair_tests.log 2>&1
cat gpurun_out/pair_tests.log
timeout 600 python bench.py --no-cpu --steps 3 --warmup 2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['e2e']['phases_last_call'], d['clocks'])"
timeout 600 python bench.py --no-cpu --steps 3 --warmup 2 --workload cfg4 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['e2e']['phases_last_call'], d['clocks'])"
timeout 600 python bench.py --no-cpu --steps 3 --warmup 2 --workload cfg2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['e2e']['phases_last_call'], d['clocks'])"
