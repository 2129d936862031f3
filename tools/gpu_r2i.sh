# round 2, run I: wider families / priors, naive mode, param_list: new tests, then the whole suite
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_wider.py -q -m gpu 2>&1 | tail -30 ) > gpurun_out/r2i_wider.log 2>&1
tail -30 gpurun_out/r2i_wider.log
( timeout 2400 python -m pytest tests -q -m gpu 2>&1 | tail -8 ) > gpurun_out/r2i_tests.log 2>&1
tail -8 gpurun_out/r2i_tests.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r2i_bench_default.log 2>&1; tail -1 gpurun_out/r2i_bench_default.log | cut -c1-2500
