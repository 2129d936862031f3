mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 ) > gpurun_out/final_tests.log 2>&1
cat gpurun_out/final_tests.log
timeout 600 python bench.py --no-cpu --steps 3 --warmup 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['e2e']['phases_last_call'], d['roofline']['frac'], d['roofline']['dram_gbs_from_traffic'])"
