mkdir -p gpurun_out
unset CGG_PROFILE
CMD2="python bench.py --workload cfg3 --cols 100 --steps 1 --warmup 1 --burnin-iters 30 --no-e2e --no-cpu"
$CMD2 > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sweep_persistent -s 2 -c 1 -f -o gpurun_out/prof_binom_stationary $CMD2 > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log | cut -c1-200
ls -la gpurun_out/prof_binom_stationary.ncu-rep
