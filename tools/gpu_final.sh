# the round-end sequence on one GPU: smoke, the whole GPU suite, the default bench and the reference arm
mkdir -p gpurun_out
python __graft_entry__.py --smoke > gpurun_out/final_smoke.log 2>&1; tail -1 gpurun_out/final_smoke.log
( timeout 2400 python -m pytest tests -q -m gpu 2>&1 | tail -6 ) > gpurun_out/final_tests.log 2>&1; tail -6 gpurun_out/final_tests.log
python bench.py > gpurun_out/final_bench.log 2>&1; tail -1 gpurun_out/final_bench.log | cut -c1-400
python bench.py --impl reference > gpurun_out/final_bench_ref.log 2>&1; tail -1 gpurun_out/final_bench_ref.log | cut -c1-400
