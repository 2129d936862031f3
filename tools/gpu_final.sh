# final validation of a round: full GPU suite, smoke, default bench (with e2e + cpu baseline), reference arm
mkdir -p gpurun_out
unset CGG_PROFILE
( timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) > gpurun_out/final_tests.log 2>&1
cat gpurun_out/final_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; tail -1 gpurun_out/final_smoke.log
timeout 900 python bench.py > gpurun_out/final_bench.log 2>&1; tail -1 gpurun_out/final_bench.log | cut -c1-3000
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_ref.log 2>&1; tail -1 gpurun_out/final_bench_ref.log | cut -c1-300
