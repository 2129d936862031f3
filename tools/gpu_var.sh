mkdir -p gpurun_out
B="python bench.py --no-cpu --steps 3 --warmup 2"
P='import sys,json; d=json.loads(sys.stdin.read()); print(round(d["value"]), round(d["e2e"]["value"]), d["e2e"]["phases_last_call"]["run_ms"], d["clocks"]["sm_mhz"])'
( echo "== 384"; timeout 300 $B 2>&1 | tail -1 | python -c "$P"
echo "== 256"; CGG_LIB=$PWD/tools/var/lib_256.so timeout 300 $B 2>&1 | tail -1 | python -c "$P"
echo "== 256p2"; CGG_LIB=$PWD/tools/var/lib_256p2.so timeout 300 $B 2>&1 | tail -1 | python -c "$P"
echo "== 256p2 cfg4"; CGG_LIB=$PWD/tools/var/lib_256p2.so timeout 300 $B --workload cfg4 --no-e2e 2>&1 | tail -1 | cut -c1-100
echo "== 256 cfg4"; CGG_LIB=$PWD/tools/var/lib_256.so timeout 300 $B --workload cfg4 --no-e2e 2>&1 | tail -1 | cut -c1-100
echo "== 384 cfg4"; timeout 300 $B --workload cfg4 --no-e2e 2>&1 | tail -1 | cut -c1-100 ) > gpurun_out/var.log 2>&1
cat gpurun_out/var.log
