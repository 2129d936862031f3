mkdir -p gpurun_out
B="python bench.py --no-cpu --steps 3 --warmup 2"
( echo "== prev"; CGG_LIB=$PWD/tools/var/lib_prev.so timeout 400 $B 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['e2e']['phases_last_call'])"
echo "== new"; timeout 400 $B 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['e2e']['phases_last_call'])"
echo "== prev cfg4"; CGG_LIB=$PWD/tools/var/lib_prev.so timeout 400 $B --workload cfg4 --no-e2e 2>&1 | tail -1 | cut -c1-100
echo "== new cfg4"; timeout 400 $B --workload cfg4 --no-e2e 2>&1 | tail -1 | cut -c1-100
echo "== new cfg2"; timeout 400 $B --workload cfg2 --no-e2e 2>&1 | tail -1 | cut -c1-100 ) > gpurun_out/var.log 2>&1
cat gpurun_out/var.log
