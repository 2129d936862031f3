mkdir -p gpurun_out
( timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/gpu_tests.log 2>&1
cat gpurun_out/gpu_tests.log
unset CGG_PROFILE
CMD2="python bench.py --workload cfg3 --cols 100 --steps 1 --warmup 1 --burnin-iters 30 --no-e2e --no-cpu"
$CMD2 > gpurun_out/plain_p100.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep_persistent -s 2 -c 1 -f -o gpurun_out/prof_binom_light $CMD2 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
