# round 2, run G (1 GPU): cluster driver fix, speculative prefetch, mailbox exchange (two shards on one device): tests + bench
mkdir -p gpurun_out
( timeout 2400 python -m pytest tests -q -m gpu -x 2>&1 | tail -25 ) > gpurun_out/r2g_tests.log 2>&1
tail -25 gpurun_out/r2g_tests.log
B="python bench.py --no-e2e --no-cpu --steps 3 --warmup 3"
P100="--workload cfg3 --cols 100"
one() { echo "== $1"; shift; timeout 600 "$@" 2>&1 | tail -1; }
( one "cfg3 p=100" $B $P100
  one "cfg3 full" $B
  one "cfg4 p=100" $B --workload cfg4 --cols 100
  one "gauss p=100" $B $P100 --family gaussian
  one "cfg2 (cluster auto)" $B --workload cfg2
  CGG_CLUSTER=8 one "cfg2 cluster=8" $B --workload cfg2
  CGG_CLUSTER=4 one "cfg2 cluster=4" $B --workload cfg2
  CGG_SMALLN=0 one "cfg2 grid" $B --workload cfg2
  one "cfg1-like n=1000 p=3 gaussian 1 chain" $B --workload cfg2 --rows 1000 --cols 3 --family gaussian --chains 1 --steps 200
  one "tiny 4 chains" $B --workload tiny --steps 20 ) > gpurun_out/r2g_bench.log 2>&1
( export CGG_PROFILE=1; echo "== profile cfg3 p=100"; timeout 300 $B $P100 2>&1 | grep "cgg profile" | tail -3 ) >> gpurun_out/r2g_bench.log 2>&1
python - <<'PY'
import json
for line in open('gpurun_out/r2g_bench.log'):
    if line.startswith('==') or line.startswith('[cgg'): print(line.strip()[:400]); continue
    try: d = json.loads(line)
    except Exception: print(line[:300]); continue
    print('   ', round(d['value']), 'upd/s', round(d['ms_per_step'], 3), 'ms/step grid', d['roofline']['grid'], 'frac', round(d['roofline']['frac'], 3), 'fallbacks', d['engine_stats']['jet_fallbacks_per_update'])
PY
